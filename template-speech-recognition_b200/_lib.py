"""ctypes binding of libtsr_b200.so (include/tsr_b200.h).  No fallback: if the library is
missing or the device is not a B200-class GPU the product path raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TSR_B200_LIB") or os.path.join(HERE, "libtsr_b200.so")   # override: debug builds (tools/trace_*.py)

TSR_OK, TSR_ERR_INVALID, TSR_ERR_CUDA, TSR_ERR_UNSUPPORTED = 0, -1, -2, -3
IMPL_TENSOR, IMPL_SIMT, IMPL_TENSOR_INT8 = 0, 1, 2
PLANES_AUTO, PLANES_BF16, PLANES_FP16, PLANES_INT8X2, PLANES_INT8X3 = 0, 1, 2, 3, 4

_vp, _i, _ll, _sz, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_size_t, ctypes.c_float

# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against the header
SIGNATURES = {
    "tsr_abi_version": (_i, []),
    "tsr_last_error": (ctypes.c_char_p, []),
    "tsr_device_supported": (_i, []),
    "tsr_rowwords": (_i, [_i]),
    "tsr_pack_edges": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "tsr_unpack_edges": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "tsr_host_pack_edges": (_i, [_vp, _vp, _ll, _i, _i]),
    "tsr_pack_features_u8": (_i, [_vp, _vp, _ll, _i, _i, _vp]),
    "tsr_unpack_features_u8": (_i, [_vp, _vp, _ll, _i, _i, _vp]),
    "tsr_spread_edges": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _vp]),
    "tsr_parts_blob_bytes": (_sz, [_i, _i, _i]),
    "tsr_prepare_parts": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "tsr_code_parts": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "tsr_code_parts_mag_aux_terms": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "tsr_patch_counts": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "tsr_patch_select_workspace_bytes": (_sz, [_ll]),
    "tsr_patch_select": (_i, [_vp, _ll, ctypes.c_uint, ctypes.c_uint, _vp, _vp, _vp, _vp]),
    "tsr_patch_gather": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "tsr_spread_codes": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "tsr_spread_graph": (_i, [_vp, _vp, _ll, _i, _i, _i, _vp]),
    "tsr_filter_blob_bytes": (_sz, [_i, _i, _i, _i]),
    "tsr_prepare_filters": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "tsr_detect": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _i, _vp]),
    "tsr_detect_reduce": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "tsr_detect_best": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tsr_detect_float": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp]),
    "tsr_detect_peaks": (_i, [_vp, _vp, _i, _i, _i, _f, _vp, _vp]),
    "tsr_peaks_gather": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "tsr_cluster_peaks": (_i, [_vp, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "tsr_edge_map": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "tsr_threshold_edges": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, ctypes.c_double, _vp, _vp]),
    "tsr_edges_from_spectrogram": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, ctypes.c_double, _vp, _vp]),
    "tsr_em_workspace_bytes": (_sz, [_ll, _i, _i]),
    "tsr_em_workspace_init": (_i, [_vp, _ll, _i, _i, _vp]),
    "tsr_em_npad": (_ll, [_ll]),
    "tsr_transpose_bits": (_i, [_vp, _ll, _i, _i, _vp, _ll, _vp]),
    "tsr_em_tiled_words": (_sz, [_ll, _i]),
    "tsr_em_tile_bits": (_i, [_vp, _ll, _i, _vp, _vp]),
    "tsr_em_estep": (_i, [_vp, _vp, _ll, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _i, _vp]),
    "tsr_em_mstep": (_i, [_vp, _vp, _vp, _ll, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "tsr_em_finalize": (_i, [_vp, _ll, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _vp, _vp]),
    "tsr_em_control_bytes": (_sz, [_i]),
    "tsr_em_control_init": (_i, [_vp, _i, ctypes.c_double, _i, _i, _i, _vp]),
    "tsr_em_control_step": (_i, [_vp, _vp, _i, _i, _vp]),
    "tsr_em_peer_flag_bytes": (_sz, [_i]),
    "tsr_em_peer_epoch_bytes": (_sz, []),
    "tsr_em_allreduce_peer": (_i, [_vp, _i, _i, _ll, _ll, _ll, _vp, _ll, _vp, _vp, _vp]),
}

ABI_VERSION = 2
_lib = None


class TsrError(RuntimeError):
    pass


def lib():
    """Load the CUDA library; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TsrError("libtsr_b200.so is not built (%s). Run `python __graft_entry__.py build`; "
                           "there is no CPU fallback." % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.tsr_abi_version() != ABI_VERSION:
            raise TsrError("libtsr_b200.so ABI version mismatch")
        _lib = l
    return _lib


def check(rc):
    if rc == TSR_OK:
        return
    msg = lib().tsr_last_error().decode("utf-8", "replace")
    if rc == TSR_ERR_INVALID:
        raise ValueError(msg)
    if rc == TSR_ERR_UNSUPPORTED:
        raise TsrError("unsupported: " + msg)
    raise TsrError("CUDA error: " + msg)


def call(name, *args):
    check(getattr(lib(), name)(*args))
