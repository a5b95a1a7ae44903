"""Device-resident batched API over the C ABI (torch tensors in, torch tensors out).

This is the layer the benchmark and the reference-signature modules (code_parts.py,
spread_waliji_patches.py, compute_likelihood_linear_filter.py, bernoulli_mixture.py)
sit on.  torch is used for device memory and streams only; every computation is a call
into libtsr_b200.so on the current CUDA stream.  Nothing here touches the CPU oracle.
"""
import os

import numpy as np
import torch

from . import _lib
from ._lib import (IMPL_SIMT, IMPL_TENSOR, IMPL_TENSOR_INT8, PLANES_AUTO, PLANES_BF16, PLANES_FP16,  # noqa: F401
                   PLANES_INT8X2, PLANES_INT8X3)

# Orientation table produced by the reference edge extractor (edge_signal_proc.py:1751-1821,
# returned by _edge_map_no_threshold and passed to spread_edgemap by every pipeline caller):
# rows are (frequency step, time step) of orientation block e.
DEFAULT_EDGE_ORIENTATIONS = np.array([[0, 1], [0, -1], [1, 0], [-1, 0], [1, 1], [-1, -1], [-1, 1], [1, -1]], np.int8)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _device_of(args, kwargs):
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.Tensor) and a.is_cuda:
            return a.device
        blob = getattr(a, "blob", None)          # PartsModel / FilterBank
        if isinstance(blob, torch.Tensor) and blob.is_cuda:
            return blob.device
    return None


def _on_device(fn):
    """Run `fn` with the CUDA device of its first device argument current: the C ABI launches on the current device
    and on torch's current stream OF THAT DEVICE, so tensors living on cuda:N must not be processed while another
    device is current.  Tensors on different devices in one call are refused."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = _device_of(args, kwargs)
        if dev is None:
            return fn(*args, **kwargs)
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda and a.device != dev:
                raise ValueError("%s: arguments live on different CUDA devices (%s, %s)" % (fn.__name__, dev, a.device))
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapped


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _dev(t, dtype, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise ValueError("%s must be a CUDA tensor" % name)
    if t.dtype != dtype:
        raise ValueError("%s: dtype mismatch, expected %s but got %s" % (name, dtype, t.dtype))
    return t.contiguous()


def rowwords(D):
    return _lib.lib().tsr_rowwords(int(D))


# ------------------------------------------------------------------------------- K0
def pack_edges(X, out=None):
    """X uint8 (B,T,F,E) -> bits uint8 (B,T,F); any nonzero counts as an edge.  out: write into this contiguous (B,T,F) tensor."""
    X = _dev(X, torch.uint8, "X")
    B, T, F, E = X.shape
    bits = torch.empty((B, T, F), dtype=torch.uint8, device=X.device) if out is None else _dev(out, torch.uint8, "out")
    if tuple(bits.shape) != (B, T, F):
        raise ValueError("out must have shape (B,T,F)")
    _lib.call("tsr_pack_edges", _ptr(X), _ptr(bits), B, T, F, E, _stream())
    return bits


def unpack_edges(bits, E=8):
    bits = _dev(bits, torch.uint8, "bits")
    B, T, F = bits.shape
    X = torch.empty((B, T, F, E), dtype=torch.uint8, device=bits.device)
    _lib.call("tsr_unpack_edges", _ptr(bits), _ptr(X), B, T, F, E, _stream())
    return X


def pack_features(Fm):
    """Fm uint8 (B,T,D) -> featbits int32 (B,T,rowwords(D)) (bit pattern of uint32 words)."""
    Fm = _dev(Fm, torch.uint8, "F")
    B, T, D = Fm.shape
    rw = rowwords(D)
    bits = torch.empty((B, T, rw), dtype=torch.int32, device=Fm.device)
    _lib.call("tsr_pack_features_u8", _ptr(Fm), _ptr(bits), B * T, D, rw, _stream())
    return bits


def unpack_features(featbits, D):
    featbits = _dev(featbits, torch.int32, "featbits")
    B, T, rw = featbits.shape
    Fm = torch.empty((B, T, D), dtype=torch.uint8, device=featbits.device)
    _lib.call("tsr_unpack_features_u8", _ptr(featbits), _ptr(Fm), B * T, D, rw, _stream())
    return Fm


# ------------------------------------------------------------------------------ K1a
def spread_edges(bits, radius, orientations=None, mode="line", block_len=0, tail_window=False):
    """Edge spreading on packed edge bytes (edge_signal_proc.py:1397-1417 / :1623-1632), inside block_len-frame blocks
    when block_len > 0.  tail_window: a ragged last block is spread as the tail of the window of the final block_len
    frames (what _edge_map_threshold_segments does, :1595-1633) instead of as a short block of its own."""
    bits = _dev(bits, torch.uint8, "bits")
    B, T, F = bits.shape
    out = torch.empty_like(bits)
    if mode == "line":
        dirs = np.ascontiguousarray(DEFAULT_EDGE_ORIENTATIONS if orientations is None else orientations)
        if dirs.shape != (8, 2) or np.any(dirs != np.round(dirs)):
            raise ValueError("orientations must be 8 integer (frequency step, time step) pairs")
        dirs = np.ascontiguousarray(dirs, dtype=np.int8)
        _lib.call("tsr_spread_edges", _ptr(bits), _ptr(out), B, T, F, int(radius), 0, dirs.ctypes.data, int(block_len),
                  int(bool(tail_window)), _stream())
    elif mode == "box":
        _lib.call("tsr_spread_edges", _ptr(bits), _ptr(out), B, T, F, int(radius), 1, 0, int(block_len), int(bool(tail_window)),
                  _stream())
    else:
        raise ValueError("mode must be 'line' or 'box'")
    return out


# ------------------------------------------------------------------------- N4 / N3
# edge_signal_proc.py:1187: per-orientation absolute floor of the quantile threshold
DEFAULT_ABST_THRESHOLD = np.array([0.025, 0.025, 0.015, 0.015, 0.02, 0.02, 0.02, 0.02], np.float64)


def edge_map(S):
    """S float64 (B,F,T) spectrograms -> raw edge magnitudes float64 (B,8,F,T)
    (edge_signal_proc._edge_map_no_threshold, :1697-1836)."""
    S = _dev(S, torch.float64, "S")
    B, F, T = S.shape
    mags = torch.empty((B, 8, F, T), dtype=torch.float64, device=S.device)
    _lib.call("tsr_edge_map", _ptr(S), _ptr(mags), B, F, T, _stream())
    return mags


def _threshold(entry, src, B, F, T, block_len, quantile, abst_threshold, want_taus):
    abst = np.ascontiguousarray(DEFAULT_ABST_THRESHOLD if abst_threshold is None else abst_threshold, dtype=np.float64)
    if abst.shape != (8,):
        raise ValueError("abst_threshold must hold 8 values (one per orientation)")
    block_len = int(block_len)
    if block_len < 1:
        raise ValueError("block_len must be positive")
    bits = torch.empty((B, T, F), dtype=torch.uint8, device=src.device)
    ragged = T % block_len != 0
    tail = torch.empty((B, min(T, block_len), F), dtype=torch.uint8, device=src.device) if ragged else None
    taus = torch.empty((B, -(-T // block_len), 8), dtype=torch.float64, device=src.device) if want_taus else None
    _lib.call(entry, _ptr(src), _ptr(bits), _ptr(tail), _ptr(taus), B, F, T, block_len, float(quantile), abst.ctypes.data,
              _stream())
    return bits, tail, taus


def threshold_edges(mags, block_len, quantile=.7, abst_threshold=None, want_taus=False):
    """Per (block, orientation) quantile thresholding of raw magnitudes (B,8,F,T) -> packed edge bytes
    (B,T,F) (threshold_edgemap per block, edge_signal_proc.py:1184-1307, 1537-1675).  Returns
    (bits, tail_bits or None, taus or None); tail_bits is the thresholded window of the final block_len
    frames when T is not a multiple of block_len (see spread_edge_blocks)."""
    mags = _dev(mags, torch.float64, "mags")
    B, E, F, T = mags.shape
    if E != 8:
        raise ValueError("mags must have 8 orientation planes")
    return _threshold("tsr_threshold_edges", mags, B, F, T, block_len, quantile, abst_threshold, want_taus)


def edges_from_spectrogram(S, block_len, quantile=.7, abst_threshold=None, want_taus=False):
    """edge_map + threshold_edges in one kernel: magnitudes are recomputed from a shared-memory tile of the
    spectrogram and never written to HBM."""
    S = _dev(S, torch.float64, "S")
    B, F, T = S.shape
    return _threshold("tsr_edges_from_spectrogram", S, B, F, T, block_len, quantile, abst_threshold, want_taus)


def spread_edge_blocks(bits, tail_bits, radius, block_len, orientations=None, mode="line"):
    """Spreading as _edge_map_threshold_segments applies it (:1617-1673): inside each block_len-frame block,
    and for a ragged last block inside the window of the final block_len frames, of which the last
    T % block_len frames are kept (:1633)."""
    if radius <= 0:
        return bits
    out = spread_edges(bits, radius, orientations, mode, block_len)
    T = bits.shape[1]
    last = T % int(block_len)
    if last:
        if tail_bits is None:
            raise ValueError("a ragged last block needs the tail window from threshold_edges")
        out[:, T - last:] = spread_edges(tail_bits, radius, orientations, mode, 0)[:, -last:]
    return out


# ------------------------------------------------------------------------------- K2
class PartsModel:
    """Device copy of a part dictionary in the coding kernel's layout (tsr_prepare_parts)."""

    def __init__(self, log_parts, log_invparts, mask=None, device="cuda", planes=PLANES_AUTO):
        lp = np.ascontiguousarray(log_parts)
        lip = np.ascontiguousarray(log_invparts)
        if lp.dtype != np.float64 or lip.dtype != np.float64:
            raise ValueError("Buffer dtype mismatch, expected 'DTYPE_t' (float64) but got '%s'" % lp.dtype)
        if lp.ndim != 4 or lip.shape != lp.shape:
            raise ValueError("Buffer has wrong number of dimensions (expected 4, got %d)" % lp.ndim)
        self.K, self.ph, self.pw, self.E = lp.shape
        nbytes = _lib.lib().tsr_parts_blob_bytes(self.K, self.ph, self.pw)
        self.blob = torch.empty(nbytes, dtype=torch.uint8, device=device)
        d_lp = torch.from_numpy(lp).to(self.blob.device)
        d_lip = torch.from_numpy(lip).to(self.blob.device)
        d_mask = None
        if mask is not None:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            if m.shape != (self.ph, self.pw):
                raise ValueError("mask must have the part window shape (%d,%d)" % (self.ph, self.pw))
            d_mask = torch.from_numpy((m != 0).astype(np.uint8)).to(self.blob.device)
        with torch.cuda.device(self.blob.device):
            _lib.call("tsr_prepare_parts", _ptr(d_lp), _ptr(d_lip), self.K, self.ph, self.pw, self.E, _ptr(d_mask),
                      int(planes), _ptr(self.blob), _stream())
            torch.cuda.current_stream().synchronize()   # d_lp/d_lip may be freed after this returns


def code_parts(bits, parts, threshold, mask_threshold=-1, want_codes=True, want_scores=False, want_counts=False,
               impl=IMPL_TENSOR, edge_spread_radius=0, edge_block_len=0):
    """bits uint8 (B,T,F) -> (codes uint8 (B,To,Fo) | None, scores float32 (B,To,Fo,K) | None,
    counts int16 (B,To,Fo) | None).  edge_spread_radius=1: `bits` is the UNSPREAD edge map and the radius-1 line
    spreading per edge_block_len-frame block (K1a, reference direction table, ragged last block as the tail of the
    final window) happens inside the kernel's producers -- no separate spread_edges launch.
    impl: IMPL_TENSOR (16-bit hi+lo weight planes, kind::f16: codes may differ from the reference at near-ties below 1e-5
    relative, scores carry 22 bits), IMPL_TENSOR_INT8 (kind::i8 + float64 refinement: codes EQUAL the reference's arg-max
    at every location, scores carry 16-bit fixed point; the faster form above 126 parts), IMPL_SIMT (fp32 cross-check)."""
    bits = _dev(bits, torch.uint8, "bits")
    B, T, F = bits.shape
    To, Fo = T - parts.ph + 1, F - parts.pw + 1
    if To < 1 or Fo < 1:
        raise ValueError("input (%d,%d) smaller than the part window (%d,%d)" % (T, F, parts.ph, parts.pw))
    dev = bits.device
    codes = torch.empty((B, To, Fo), dtype=torch.uint8, device=dev) if want_codes else None
    scores = torch.empty((B, To, Fo, parts.K), dtype=torch.float32, device=dev) if want_scores else None
    counts = torch.empty((B, To, Fo), dtype=torch.int16, device=dev) if want_counts else None
    _lib.call("tsr_code_parts", _ptr(bits), _ptr(parts.blob), parts.K, parts.ph, parts.pw, int(threshold),
              int(mask_threshold), _ptr(codes), _ptr(scores), _ptr(counts), B, T, F, int(edge_spread_radius),
              int(edge_block_len), int(impl), _stream())
    return codes, scores, counts


def code_parts_mag_aux_terms(scores, X_mags, X_aux, thresh_mask, log_mag_parts, log_mag_invparts, log_aux_parts,
                             log_aux_invparts, ph, pw):
    """In place on scores float32 (B,To,Fo,K) (edge log-likelihoods from code_parts): adds the magnitude
    and auxiliary terms of code_parts_mag_aux (code_parts.pyx:195-239) at gated locations, zeroes the
    rest.  Returns the gate uint8 (B,To,Fo)."""
    scores = _dev(scores, torch.float32, "scores")
    X_mags, X_aux, thresh_mask = (_dev(t, torch.uint8, n) for t, n in ((X_mags, "X_mags"), (X_aux, "X_aux"), (thresh_mask, "X_thresh_mask")))
    tabs = [_dev(t, torch.float64, "log tables") for t in (log_mag_parts, log_mag_invparts, log_aux_parts, log_aux_invparts)]
    B, To, Fo, K = scores.shape
    T, F = To + ph - 1, Fo + pw - 1
    mag_dim, aux_dim = X_mags.shape[2], X_aux.shape[2]
    if tuple(thresh_mask.shape) != (B, T, F) or X_mags.shape[:2] != (B, T) or X_aux.shape[:2] != (B, T):
        raise ValueError("X_mags / X_aux / X_thresh_mask do not match the edge map shape")
    if tuple(tabs[0].shape) != (K, ph + mag_dim) or tabs[1].shape != tabs[0].shape:
        raise ValueError("log_mag_parts must be (K, part_x_dim + mag_dim)")
    if tuple(tabs[2].shape) != (K, ph * aux_dim) or tabs[3].shape != tabs[2].shape:
        raise ValueError("log_aux_parts must be (K, part_x_dim * aux_dim)")
    gate = torch.empty((B, To, Fo), dtype=torch.uint8, device=scores.device)
    _lib.call("tsr_code_parts_mag_aux_terms", _ptr(scores), _ptr(gate), _ptr(X_mags), _ptr(X_aux), _ptr(thresh_mask),
              _ptr(tabs[0]), _ptr(tabs[1]), _ptr(tabs[2]), _ptr(tabs[3]), B, T, F, K, int(ph), int(pw), mag_dim, aux_dim,
              _stream())
    return gate


# ------------------------------------------------------------------------------- N1
def patch_counts(bits, ph, pw, mask=None):
    """bits uint8 (B,T,F) -> int32 (B,To,Fo) edge counts of every ph x pw window (under the 0/1 window mask)."""
    bits = _dev(bits, torch.uint8, "bits")
    B, T, F = bits.shape
    if T < ph or F < pw:
        raise ValueError("input (%d,%d) smaller than the window (%d,%d)" % (T, F, ph, pw))
    if mask is not None:
        mask = _dev(mask, torch.uint8, "mask")
        if tuple(mask.shape) != (ph, pw):
            raise ValueError("mask must be (%d,%d)" % (ph, pw))
    counts = torch.empty((B, T - ph + 1, F - pw + 1), dtype=torch.int32, device=bits.device)
    _lib.call("tsr_patch_counts", _ptr(bits), _ptr(mask), B, T, F, int(ph), int(pw), _ptr(counts), _stream())
    return counts


def extract_patches(bits, ph, pw, lo, hi=None, mask=None, S=None, E=8, want_bytes=False, want_bits=True):
    """Ordered compaction of the windows with lo <= count < hi (code_parts.pyx:316-347).
    Returns dict(counts, n, locs (n,3) int32 (b,i,j), bits (n,rowwords) int32 | None,
    patches (n,ph,pw,E) uint8 | None, S_patches (n,ph,pw) float64 | None).  One device->host read
    of the patch count sizes the outputs."""
    counts = patch_counts(bits, ph, pw, mask)
    B, T, F = bits.shape
    n_loc = counts.numel()
    dev = bits.device
    slot = torch.empty(n_loc, dtype=torch.int32, device=dev)
    total = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = torch.empty(max(_lib.lib().tsr_patch_select_workspace_bytes(n_loc), 4), dtype=torch.uint8, device=dev)
    lo = max(0, min(int(lo), 2 ** 32 - 1))
    hi = 2 ** 32 - 1 if hi is None else max(0, min(int(hi), 2 ** 32 - 1))
    _lib.call("tsr_patch_select", _ptr(counts), n_loc, lo, hi, _ptr(slot), _ptr(total), _ptr(ws), _stream())
    n = int(total.item())
    rw = rowwords(ph * pw * E)
    locs = torch.empty((n, 3), dtype=torch.int32, device=dev)
    pbits = torch.empty((n, rw), dtype=torch.int32, device=dev) if want_bits else None
    pset = torch.empty((n, ph, pw, E), dtype=torch.uint8, device=dev) if want_bytes else None
    Sp = None
    if S is not None:
        S = _dev(S, torch.float64, "S")
        if tuple(S.shape) != (B, T, F):
            raise ValueError("S must match the edge map shape")
        Sp = torch.empty((n, ph, pw), dtype=torch.float64, device=dev)
    if n:
        _lib.call("tsr_patch_gather", _ptr(bits), _ptr(S), _ptr(slot), B, T, F, int(E), int(ph), int(pw), _ptr(pset), _ptr(Sp),
                  _ptr(locs), _ptr(pbits), _stream())
    return dict(counts=counts, n=n, locs=locs, bits=pbits, patches=pset, S_patches=Sp)


# ------------------------------------------------------------------------------ K1b
def spread_codes(codes, K, r0, r1):
    """codes uint8 (B,To,Fo) -> featbits int32 (B,To,rowwords(Fo*K))."""
    codes = _dev(codes, torch.uint8, "codes")
    B, To, Fo = codes.shape
    rw = rowwords(Fo * K)
    out = torch.empty((B, To, rw), dtype=torch.int32, device=codes.device)
    _lib.call("tsr_spread_codes", _ptr(codes), _ptr(out), B, To, Fo, int(K), int(r0), int(r1), rw, _stream())
    return out


def spread_graph(featbits, graph, Fo, K):
    """In place on featbits (B,To,rowwords): part-graph spreading (code_parts.pyx:459-503)."""
    featbits = _dev(featbits, torch.int32, "featbits")
    graph = _dev(graph, torch.uint8, "graph")
    if tuple(graph.shape) != (K, K):
        raise ValueError("part_graph must be (%d,%d)" % (K, K))
    R = featbits.shape[0] * featbits.shape[1]
    _lib.call("tsr_spread_graph", _ptr(featbits), _ptr(graph), R, int(Fo), int(K), featbits.shape[2], _stream())
    return featbits


# ------------------------------------------------------------------------------- K3
class FilterBank:
    """M linear filters of L frames x D features (compute_likelihood_linear_filter's LF),
    resident on the device both as fp32 (SIMT cross-check) and as the tensor kernel's blob."""

    def __init__(self, LF, cs=None, device="cuda", planes=PLANES_AUTO):
        LF = np.ascontiguousarray(LF)
        if LF.dtype != np.float32:
            raise ValueError("Buffer dtype mismatch, expected 'DTYPE_t' (float32) but got '%s'" % LF.dtype)
        if LF.ndim < 3:
            raise ValueError("LF must be (M, L, ...)")
        self.M, self.L = LF.shape[0], LF.shape[1]
        self.D = int(np.prod(LF.shape[2:]))
        self.LF = torch.from_numpy(LF.reshape(self.M, self.L, self.D)).to(device)
        self.cs = None if cs is None else torch.as_tensor(np.asarray(cs, dtype=np.float32).reshape(self.M)).to(device)
        self.planes = int(planes)
        nbytes = _lib.lib().tsr_filter_blob_bytes(self.M, self.L, self.D, self.planes)
        if nbytes == 0:
            raise ValueError("unsupported filter shape M=%d L=%d D=%d (need 1 <= L <= 256)" % (self.M, self.L, self.D))
        self.blob = torch.empty(nbytes, dtype=torch.uint8, device=device)
        with torch.cuda.device(self.blob.device):
            _lib.call("tsr_prepare_filters", _ptr(self.LF), self.M, self.L, self.D, int(planes), _ptr(self.blob), _stream())


def detect(featbits, bank, D=None, impl=IMPL_TENSOR, out=None):
    """featbits int32 (B,T,rowwords) -> scores float32 (B,T,M), zero in the L-1 tail frames."""
    featbits = _dev(featbits, torch.int32, "featbits")
    B, T, rw = featbits.shape
    D = bank.D if D is None else D
    if rw != rowwords(D) or D != bank.D:
        raise ValueError("feature width %d does not match the filters (%d)" % (D, bank.D))
    if out is None:
        out = torch.empty((B, T, bank.M), dtype=torch.float32, device=featbits.device)
    _lib.call("tsr_detect", _ptr(featbits), B, T, D, rw, _ptr(bank.blob), _ptr(bank.LF), bank.M, bank.L, bank.planes,
              _ptr(out), int(impl), _stream())
    return out


def detect_best(featbits, bank, D=None, want_scores=False):
    """detect + the mixture max fused into the kernel epilogue: featbits int32 (B,T,rowwords) ->
    (best float32 (B,T), arg int16 (B,T) holding uint16 ids, scores (B,T,M) | None).  Same bits as
    detect() followed by detect_reduce(..., tail_mode=1), without the (B,T,M) score matrix."""
    featbits = _dev(featbits, torch.int32, "featbits")
    B, T, rw = featbits.shape
    D = bank.D if D is None else D
    if rw != rowwords(D) or D != bank.D:
        raise ValueError("feature width %d does not match the filters (%d)" % (D, bank.D))
    if bank.L > T:
        raise ValueError("templates of %d frames do not fit utterances of %d" % (bank.L, T))
    dev = featbits.device
    best = torch.empty((B, T), dtype=torch.float32, device=dev)
    arg = torch.empty((B, T), dtype=torch.int16, device=dev)
    keys = torch.empty((B, T), dtype=torch.int64, device=dev)
    scores = torch.empty((B, T, bank.M), dtype=torch.float32, device=dev) if want_scores else None
    _lib.call("tsr_detect_best", _ptr(featbits), B, T, D, rw, _ptr(bank.blob), bank.M, bank.L, bank.planes, _ptr(bank.cs),
              _ptr(scores), _ptr(keys), _ptr(best), _ptr(arg), _stream())
    return best, arg, scores


def detect_reduce(scores, cs, L, tail_mode=1):
    """scores (B,T,M), cs (M,) | None -> (best float32 (B,T), arg int16 (B,T) holding uint16 ids)."""
    scores = _dev(scores, torch.float32, "scores")
    B, T, M = scores.shape
    best = torch.empty((B, T), dtype=torch.float32, device=scores.device)
    arg = torch.empty((B, T), dtype=torch.int16, device=scores.device)
    _lib.call("tsr_detect_reduce", _ptr(scores), _ptr(cs), B, T, M, int(L), int(tail_mode), _ptr(best), _ptr(arg), _stream())
    return best, arg


def detect_float(Fm, LF, init=None):
    """Fm float32 (B,T,D), LF float32 (M,L,D) device tensors -> scores (B,T,M); accumulations
    start at init[m] when given."""
    Fm = _dev(Fm, torch.float32, "F")
    LF = _dev(LF, torch.float32, "LF")
    B, T, D = Fm.shape
    M, L, D2 = LF.shape
    if D != D2:
        raise ValueError("feature width mismatch")
    out = torch.empty((B, T, M), dtype=torch.float32, device=Fm.device)
    _lib.call("tsr_detect_float", _ptr(Fm), B, T, D, _ptr(LF), M, L, _ptr(init), _ptr(out), _stream())
    return out


# ------------------------------------------------------------------------------- N2
def detect_peaks(track, threshold, C0=1, lengths=None):
    """track float32 (B,T) -> (peak_bt int32 (n,2) rows (b,t) in order, peak_val float32 (n,)): local maxima
    >= threshold over t in [1, length-C0-1] (roc_functions.py:177-186).  One device->host read of n."""
    track = _dev(track, torch.float32, "track")
    B, T = track.shape
    dev = track.device
    if lengths is not None:
        lengths = _dev(lengths, torch.int32, "lengths")
    n_loc = B * T
    flags = torch.empty(n_loc, dtype=torch.int32, device=dev)
    _lib.call("tsr_detect_peaks", _ptr(track), _ptr(lengths), B, T, int(C0), float(threshold), _ptr(flags), _stream())
    slot = torch.empty(n_loc, dtype=torch.int32, device=dev)
    total = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = torch.empty(max(_lib.lib().tsr_patch_select_workspace_bytes(n_loc), 4), dtype=torch.uint8, device=dev)
    _lib.call("tsr_patch_select", _ptr(flags), n_loc, 1, 2 ** 32 - 1, _ptr(slot), _ptr(total), _ptr(ws), _stream())
    n = int(total.item())
    bt = torch.empty((n, 2), dtype=torch.int32, device=dev)
    val = torch.empty(n, dtype=torch.float32, device=dev)
    if n:
        _lib.call("tsr_peaks_gather", _ptr(track), _ptr(slot), B, T, _ptr(bt), _ptr(val), _stream())
    return bt, val


def cluster_peaks(peak_bt, B, C0, C1):
    """Cluster the peaks of every utterance (cluster_times.pyx:133-219 + roc_functions.get_cluster_starts_ends).
    peak_bt int32 (n,2) from detect_peaks.  Returns (offsets int32 (B+1,), partition uint8 (n+B,), clusters int32
    (n,2) rows (start, end) -- utterance b's from row offsets[b] --, ncluster int32 (B,))."""
    peak_bt = _dev(peak_bt, torch.int32, "peak_bt")
    n, dev = peak_bt.shape[0], peak_bt.device
    offsets = torch.empty(B + 1, dtype=torch.int32, device=dev)
    partition = torch.empty(max(n + B, 1), dtype=torch.uint8, device=dev)
    clusters = torch.empty((n, 2), dtype=torch.int32, device=dev)
    ncluster = torch.empty(B, dtype=torch.int32, device=dev)
    _lib.call("tsr_cluster_peaks", _ptr(peak_bt), n, int(B), float(C0), float(C1), _ptr(offsets), _ptr(partition),
              _ptr(clusters), _ptr(ncluster), _stream())
    return offsets, partition[:n + B], clusters, ncluster


# ------------------------------------------------------------------------------- K4
EM_QMAX = 8355711            # posteriors are quantised to q = round(aff / s_k * EM_QMAX), s_k = 2^ceil(log2 max_n aff[n][k])


class EmControl:
    """Device-side stopping rule of the EM loop (tsr_em_control_*): lets the host queue iterations without
    reading the log-likelihood back; kernels of iterations past the stop exit at once."""

    def __init__(self, device, tol, mode, min_iter, max_iter, max_trace):
        self.max_trace = int(max_trace)
        nbytes = _lib.lib().tsr_em_control_bytes(self.max_trace)
        self.buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            _lib.call("tsr_em_control_init", _ptr(self.buf), self.max_trace, float(tol), int(mode), int(min_iter), int(max_iter),
                      _stream())

    def step(self, stats, D, K):
        with torch.cuda.device(self.buf.device):
            _lib.call("tsr_em_control_step", _ptr(self.buf), _ptr(stats), int(D), int(K), _stream())

    def read(self):
        """(done, iterations, trace) -- one device->host copy (synchronises the stream)."""
        h = self.buf.cpu().numpy()
        done, iterations, ntrace = (int(v) for v in h[:12].view(np.int32))
        trace = h[48:48 + 8 * min(ntrace, self.max_trace)].view(np.float64)
        return bool(done), iterations, [float(v) for v in trace]


class EmState:
    """Device state of one rank's share of a Bernoulli-mixture EM problem."""

    def __init__(self, Xbits, D, K):
        self.X = _dev(Xbits, torch.int32, "Xbits")
        if self.X.dim() != 2:
            raise ValueError("Xbits must be (N, rowwords)")
        self.N, self.rw = self.X.shape
        self.D, self.K = int(D), int(K)
        if self.rw != rowwords(D):
            raise ValueError("Xbits width does not match D")
        dev = self.X.device
        nbytes = _lib.lib().tsr_em_workspace_bytes(self.N, self.D, self.K)
        if nbytes == 0:
            raise ValueError("unsupported EM shape (need 1 <= num_mix <= 4095 and fewer than 2^24 rows per rank)")
        with torch.cuda.device(dev):
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.call("tsr_em_workspace_init", _ptr(self.ws), self.N, self.D, self.K, _stream())
            # one-time re-layouts for the tensor-core contractions (1 bit per datum each): X in tile order for the E-step,
            # X transposed and in tile order for the M-step
            self.Npad = int(_lib.lib().tsr_em_npad(self.N))
            self.Xt = self.XTt = None
            if self.N:
                XT = torch.zeros((self.D, self.Npad // 32), dtype=torch.int32, device=dev)
                _lib.call("tsr_transpose_bits", _ptr(self.X), self.N, self.D, self.rw, _ptr(XT), self.Npad, _stream())
                self.Xt = torch.empty(_lib.lib().tsr_em_tiled_words(self.N, self.rw), dtype=torch.int32, device=dev)
                _lib.call("tsr_em_tile_bits", _ptr(self.X), self.N, self.rw, _ptr(self.Xt), _stream())
                self.XTt = torch.empty(_lib.lib().tsr_em_tiled_words(self.D, self.Npad // 32), dtype=torch.int32, device=dev)
                _lib.call("tsr_em_tile_bits", _ptr(XT), self.D, self.Npad // 32, _ptr(self.XTt), _stream())
                torch.cuda.current_stream().synchronize()        # XT is released on return
                del XT
        self.aff = torch.zeros((self.N, self.K), dtype=torch.float32, device=dev)
        # stats = [K*D numerators | K responsibility sums | log-likelihood], float64 (see tsr_b200.h)
        self.stats = torch.zeros(self.K * self.D + self.K + 1, dtype=torch.float64, device=dev)

    def estep(self, logT, logInvT, logw, hard=False, impl=IMPL_TENSOR, scores_out=None, update=True, ctl=None, stats=None):
        """logT = logInvT = None: use the weight planes the last em_finalize(..., state=self) laid out.
        stats: write this rank's statistics there instead of self.stats (a peer-mapped source slot, see PeerStats)."""
        stats = self.stats if stats is None else stats
        with torch.cuda.device(self.X.device):
            _lib.call("tsr_em_estep", _ptr(self.X), _ptr(self.Xt), self.N, self.D, self.rw, self.K, _ptr(logT), _ptr(logInvT), _ptr(logw),
                      _ptr(self.aff) if update else 0, _ptr(scores_out), _ptr(stats) if update else 0, int(bool(hard)),
                      int(bool(update)), _ptr(self.ws), _ptr(ctl.buf) if ctl is not None else 0, int(impl), _stream())

    def mstep_accumulate(self, impl=IMPL_TENSOR, aff=None, ctl=None, stats=None):
        """stats[:K*D] = numerators from the last E-step's posteriors, or from `aff` (N,K) float32 when given
        (then the responsibility sums are taken from it too)."""
        stats = self.stats if stats is None else stats
        if aff is not None:
            aff = _dev(aff, torch.float32, "aff")
        elif impl == IMPL_SIMT:
            aff = self.aff
        with torch.cuda.device(self.X.device):
            _lib.call("tsr_em_mstep", _ptr(self.X), 0, _ptr(self.XTt), self.N, self.D, self.rw, self.K, _ptr(aff), _ptr(stats),
                      _ptr(self.ws), _ptr(ctl.buf) if ctl is not None else 0, int(impl), _stream())

    def numerators(self):
        return self.stats[:self.K * self.D].reshape(self.K, self.D)

    def resp_sums(self):
        return self.stats[self.K * self.D:self.K * self.D + self.K]

    def loglik(self):
        return float(self.stats[self.K * self.D + self.K].item())


class PeerStats:
    """Home of a rank's EM statistics vector in NVLink peer-mapped (symmetric) memory, for tsr_em_allreduce_peer:
    one allocation per rank holding two source slots (alternated from launch to launch) and the flag array, mapped
    by every rank of `group`.  Raises when symmetric memory is unavailable (the driver then keeps the NCCL allreduce)."""

    def __init__(self, n, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        group = dist.group.WORLD if group is None else group
        self.world, self.rank, self.n = dist.get_world_size(group), dist.get_rank(group), int(n)
        self.slot_bytes = (self.n * 8 + 255) // 256 * 256
        self.flag_off = 3 * self.slot_bytes                       # [source slot 0 | source slot 1 | result slot | flags]
        tp = os.environ.get("TSR_EM_PEER_TWO_PHASE")            # testing: "1" / "0" force the form
        self.two_phase = (self.world > 2) if tp is None else tp == "1"   # reduce-scatter + all-gather by remote stores
        total = self.flag_off + int(_lib.lib().tsr_em_peer_flag_bytes(self.world))
        with torch.cuda.device(device):
            self.buf = symm.empty(total, dtype=torch.uint8, device=device)
            self.buf.zero_()
            self.handle = symm.rendezvous(self.buf, group.group_name if hasattr(group, "group_name") else group)
            ptrs = [int(p) for p in self.handle.buffer_ptrs]
            if len(ptrs) != self.world or ptrs[self.rank] != self.buf.data_ptr():
                raise RuntimeError("symmetric memory rendezvous returned unexpected peer pointers")
            self.ptrs = torch.tensor(ptrs, dtype=torch.int64, device=device)
            self.epoch = torch.zeros(int(_lib.lib().tsr_em_peer_epoch_bytes()) // 4, dtype=torch.int32, device=device)
            self.src = [self.buf[i * self.slot_bytes:i * self.slot_bytes + self.n * 8].view(torch.float64) for i in (0, 1)]
            self.out = self.buf[2 * self.slot_bytes:2 * self.slot_bytes + self.n * 8].view(torch.float64)   # result slot (peers store into it)
            torch.cuda.synchronize(device)
        dist.barrier(group)                     # every rank's flags are zero before anybody signals
        self.launches = 0

    def allreduce(self, slot, out, ctl=None):
        """out (float64 [n]) = sum over ranks of their source slot `slot`; one kernel launch on the current stream.
        With more than two ranks the two-phase kernel leaves the result in self.out (the symmetric result slot): pass
        out=self.out (the driver makes the engine's statistics vector that view)."""
        two = self.two_phase
        if two and out.data_ptr() != self.out.data_ptr():
            raise ValueError("the two-phase peer allreduce writes the symmetric result slot: pass out=PeerStats.out")
        with torch.cuda.device(self.buf.device):
            _lib.call("tsr_em_allreduce_peer", _ptr(self.ptrs), self.rank, self.world, int(slot) * self.slot_bytes, self.flag_off, self.n,
                      0 if two else _ptr(out), 2 * self.slot_bytes if two else -1, _ptr(self.epoch),
                      _ptr(ctl.buf) if ctl is not None else 0, _stream())
        self.launches += 1


def em_finalize(stats, N_total, D, K, min_prob, templates, logT, logInvT, weights, logw, state=None, ctl=None):
    """M-step proper from the (all-reduced) statistics.  state: also lay out the weight planes of the next E-step
    in that EmState's workspace (then its estep may be called with logT = logInvT = None)."""
    stats = _dev(stats, torch.float64, "stats")
    with torch.cuda.device(stats.device):
        _lib.call("tsr_em_finalize", _ptr(stats), int(N_total), int(D), int(K), float(min_prob), _ptr(templates), _ptr(logT),
                  _ptr(logInvT), _ptr(weights), _ptr(logw), _ptr(state.ws) if state is not None else 0,
                  int(state.N) if state is not None else 0, _ptr(ctl.buf) if ctl is not None else 0, _stream())


# every tensor-taking entry point runs with its arguments' device current (see _on_device)
for _name in ("pack_edges", "unpack_edges", "pack_features", "unpack_features", "spread_edges", "edge_map", "threshold_edges",
              "edges_from_spectrogram", "spread_edge_blocks", "code_parts", "code_parts_mag_aux_terms", "patch_counts",
              "extract_patches", "spread_codes", "spread_graph", "detect", "detect_best", "detect_reduce", "detect_float", "detect_peaks",
              "cluster_peaks", "em_finalize"):
    globals()[_name] = _on_device(globals()[_name])
del _name
