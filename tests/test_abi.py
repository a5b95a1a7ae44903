"""C-ABI checks that need no GPU: the library loads, exports every symbol include/tsr_b200.h
declares, the ctypes table matches the header, and argument validation / the no-fallback rule
work without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "tsr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = re.findall(r"\b(?:int|size_t|long long|const char\*)\s+(tsr_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)
    return {name: [a.strip() for a in args.split(",") if a.strip() and a.strip() != "void"] for name, args in decls}


def test_library_exports_every_declared_symbol():
    from tsr_b200 import _lib
    funcs = _header_functions()
    assert len(funcs) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in funcs:
        assert hasattr(lib, name), "libtsr_b200.so does not export %s" % name
    # binding table and header agree on names and arity
    assert set(funcs) == set(_lib.SIGNATURES)
    for name, args in funcs.items():
        assert len(args) == len(_lib.SIGNATURES[name][1]), name


def test_size_helpers_and_validation_without_device():
    from tsr_b200 import _lib
    lib = _lib.lib()
    assert lib.tsr_abi_version() == 2
    assert lib.tsr_rowwords(3600) == 114 and lib.tsr_rowwords(64) == 2 and lib.tsr_rowwords(65) == 4
    assert lib.tsr_parts_blob_bytes(100, 5, 5) > 2 * 13 * 112 * 32
    assert lib.tsr_filter_blob_bytes(32, 40, 3600, 0) == 256 + 6 * 57 * 2 * 240 * 128  # header + 6 N-tiles x 57 chunks x 2 planes x 240 columns x 128 B
    assert lib.tsr_filter_blob_bytes(1, 300, 10, 0) == 0          # L > 256 unsupported
    # invalid arguments are rejected before any CUDA call
    assert lib.tsr_pack_edges(None, None, 1, 1, 1, 8, None) == _lib.TSR_ERR_INVALID
    assert b"null" in lib.tsr_last_error()
    buf = (ctypes.c_uint8 * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.tsr_pack_edges(p, p, 1, 1, 1, 9, None) == _lib.TSR_ERR_INVALID
    with pytest.raises(ValueError):
        _lib.check(lib.tsr_spread_codes(p, p, 1, 4, 4, 300, 1, 1, 2, None))


def test_no_cpu_fallback():
    """Without a CUDA device the compute entry points refuse instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tsr_b200 import _lib, code_parts
    lib = _lib.lib()
    buf = (ctypes.c_uint8 * 4096)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.tsr_pack_edges(p, p, 1, 2, 2, 8, None) == _lib.TSR_ERR_UNSUPPORTED
    assert lib.tsr_device_supported() == 0
    X = np.zeros((8, 8, 8), np.uint8)
    lp = np.zeros((2, 5, 5, 8))
    with pytest.raises(_lib.TsrError):
        code_parts.code_parts(X, lp, lp, 1)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "template-speech-recognition_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the CPU oracle", ""), "%s mentions the oracle" % f


def test_cython_style_argument_errors():
    from tsr_b200 import code_parts, compute_likelihood_linear_filter as clf
    X = np.zeros((8, 8, 8), np.uint8)
    with pytest.raises(ValueError, match="Buffer dtype mismatch"):
        code_parts.code_parts(X, np.zeros((2, 5, 5, 8), np.float32), np.zeros((2, 5, 5, 8), np.float32), 1)
    with pytest.raises(ValueError, match="wrong number of dimensions"):
        code_parts.code_parts(X[0], np.zeros((2, 5, 5, 8)), np.zeros((2, 5, 5, 8)), 1)
    with pytest.raises(ValueError, match="Buffer dtype mismatch"):
        code_parts.spread_patches(np.zeros((4, 4), np.int32), 1, 1, 3)
    with pytest.raises(ValueError, match="65535"):
        code_parts.spread_patches(np.zeros((70000, 1), np.int64), 1, 1, 3)
    with pytest.raises(ValueError, match="Buffer dtype mismatch"):
        clf.detect(np.zeros((4, 2, 2), np.uint8), np.zeros((2, 2, 2), np.float64))
    z2, z4 = np.zeros((8, 2), np.uint8), np.zeros((3, 5, 5, 8))
    with pytest.raises(ValueError, match="Buffer dtype mismatch"):   # X_mags must be uint8 (code_parts.pyx:137)
        code_parts.code_parts_mag_aux(np.zeros((8, 8, 8), np.uint8), z2.astype(np.float64), z2, z4, z4, np.zeros((3, 7)),
                                      np.zeros((3, 7)), np.zeros((3, 10)), np.zeros((3, 10)), np.zeros((8, 8), np.uint8))
    with pytest.raises(ValueError, match="wrong number of dimensions"):
        code_parts.get_parts_mask(np.zeros((8, 8), np.uint8), np.ones((5, 5), np.uint8), np.zeros((8, 8)), 3)


def test_host_pack_edges_matches_packbits():
    """tsr_host_pack_edges (the host-side boundary conversion of ScoringPipeline.stream: uint8 (..., 8) edge tensors ->
    one byte per (t, f)) against np.packbits, for every loop of its body (AVX-512 / SSE2 main loops, 8-cell and scalar
    tails), any non-zero byte counting as an edge, single- and multi-threaded."""
    import numpy as np
    from tsr_b200 import _lib
    lib = _lib.lib()
    rs = np.random.RandomState(3)
    X = (rs.rand(70001, 8) < 0.2).astype(np.uint8) * rs.randint(1, 256, size=(70001, 8)).astype(np.uint8)
    want = np.packbits(X != 0, axis=-1, bitorder="little")[:, 0]
    for n in (0, 1, 7, 8, 9, 31, 32, 33, 65, 4099, 70001):
        for off in (0, 3):                      # unaligned starts
            if off + n > X.shape[0]:
                continue
            Xs = np.ascontiguousarray(X[off:off + n])
            for nt in (1, 4):
                out = np.full(n + 1, 0xA5, np.uint8)
                rc = lib.tsr_host_pack_edges(Xs.ctypes.data if n else X.ctypes.data, out.ctypes.data, n, 8, nt)
                assert rc == 0
                np.testing.assert_array_equal(out[:n], want[off:off + n])
                assert out[n] == 0xA5            # nothing written past the end
