"""CPU oracle for the template_speech_rec scoring hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
import this module, and only as the checker or the timed CPU arm.  The product package
never imports it and fails loudly when its CUDA library is missing.

Two layers:

* ctypes wrappers around oracle/_build/libtsr_oracle.so (tsr_oracle.c), the plain-C
  restatement of the reference Cython (code_parts.pyx, spread_waliji_patches.pyx,
  compute_likelihood_linear_filter.pyx) and of edge_signal_proc._spread_edges;
* NumPy restatements of the Python-2-only pieces that cannot be imported under
  Python 3: bernoulli_mixture.BernoulliMixture (bernoulli_mixture.py:9-311),
  bernoulli_em.Bernoulli_Mixture (bernoulli_em.py:8-131),
  estimate_template.construct_linear_filter (estimate_template.py:230-251),
  test_template.score_template_background_section (test_template.py:25-41) and
  edge_signal_proc.spread_edgemap (edge_signal_proc.py:1397-1417).

Pinning status (see DESIGN.md "Oracle"):
  coding / code spreading / detection : pinned against the reference Cython compiled
      unmodified into oracle/_ref (tests/test_oracle.py) and the golden fixtures made
      from it (tests/golden/make_golden.py).
  Bernoulli mixture EM (both classes), linear filter construction, single-window score :
      pinned by tests/golden/ref_models.npz, produced by EXECUTING the reference's own classes
      and functions from their source (tests/golden/make_golden_models.py; Python-2 idioms
      patched as documented there), plus the docstring known-answer at bernoulli_mixture.py:50-71.
  edge front end (N4 edge extraction, N3 block thresholding) and edge spreading : pinned by
      tests/golden/ref_edges.npz, produced the same way (tests/golden/make_golden_edges.py).
  No part of the oracle is unpinned.
"""
import ctypes
import importlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile tsr_oracle.c (gcc, seconds).  Called by __graft_entry__.build()."""
    so = os.path.join(HERE, "_build", "libtsr_oracle.so")
    src = os.path.join(HERE, "tsr_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def load_ref():
    """Return the compiled reference modules (code_parts, spread_waliji_patches,
    compute_likelihood_linear_filter) from oracle/_ref, or None when not built."""
    ref = os.path.join(HERE, "_ref")
    if not os.path.isdir(ref) or not any(f.endswith(".so") for f in os.listdir(ref)):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        mods = tuple(importlib.import_module(m) for m in
                     ("code_parts", "spread_waliji_patches", "compute_likelihood_linear_filter"))
    except ImportError:
        return None
    for m in mods:  # make sure we did not pick up the product's drop-in modules
        if os.path.dirname(os.path.abspath(m.__file__)) != ref:
            return None
    return mods


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


_u8, _f32, _f64, _i64, _u16, _u32 = (ctypes.c_uint8, ctypes.c_float, ctypes.c_double, ctypes.c_int64,
                                     ctypes.c_uint16, ctypes.c_uint32)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


# --------------------------------------------------------------------------- coding
def _code(fn, X, lp, lip, *thr, M=None):
    X = _c(X, np.uint8); lp = _c(lp, np.float64); lip = _c(lip, np.float64)
    T, F, E = X.shape
    K, ph, pw, pe = lp.shape
    assert pe == E and lip.shape == lp.shape
    out = np.empty((T - ph + 1, F - pw + 1, K + 1), np.float64)
    args = [_p(X, _u8), T, F, E]
    if M is not None:
        M = _c(M, np.uint8)
        args.append(_p(M, _u8))
    args += [_p(lp, _f64), _p(lip, _f64), K, ph, pw] + [int(t) for t in thr] + [_p(out, _f64)]
    getattr(lib(), fn)(*args)
    return out


def code_parts(X, log_parts, log_invparts, threshold):
    """code_parts.pyx:70-135"""
    return _code("oracle_code_parts", X, log_parts, log_invparts, threshold)


def code_parts_fast(X, log_parts, log_invparts, threshold):
    """code_parts.pyx:247-295"""
    return _code("oracle_code_parts_fast", X, log_parts, log_invparts, threshold)


def code_parts_mask(X, M, log_parts, log_invparts, mask_threshold, absolute_threshold):
    """code_parts.pyx:353-404 (window-relative mask; see tsr_oracle.c note)"""
    return _code("oracle_code_parts_mask", X, log_parts, log_invparts, mask_threshold, absolute_threshold, M=M)


def code_parts_mag_aux(X_edges, X_mags, X_aux, log_parts, log_invparts, log_mag_parts, log_mag_invparts,
                       log_aux_parts, log_aux_invparts, X_thresh_mask):
    """code_parts.pyx:138-242"""
    X = _c(X_edges, np.uint8); Xm = _c(X_mags, np.uint8); Xa = _c(X_aux, np.uint8); tm = _c(X_thresh_mask, np.uint8)
    lp, lip, lmp, lmip, lap, laip = (_c(a, np.float64) for a in (log_parts, log_invparts, log_mag_parts, log_mag_invparts,
                                                                  log_aux_parts, log_aux_invparts))
    T, F, E = X.shape
    K, ph, pw, _ = lp.shape
    assert lmp.shape == (K, ph + Xm.shape[1]) and lap.shape == (K, ph * Xa.shape[1])
    out = np.empty((T - ph + 1, F - pw + 1, K + 1), np.float64)
    lib().oracle_code_parts_mag_aux(_p(X, _u8), T, F, E, _p(Xm, _u8), Xm.shape[1], _p(Xa, _u8), Xa.shape[1], _p(lp, _f64),
                                    _p(lip, _f64), _p(lmp, _f64), _p(lmip, _f64), _p(lap, _f64), _p(laip, _f64),
                                    _p(tm, _u8), K, ph, pw, _p(out, _f64))
    return out


def spread_patches(X, r0, r1, num_parts):
    """code_parts.pyx:406-457 == spread_waliji_patches.pyx:15-59"""
    X = _c(X, np.int64)
    out = np.empty(X.shape + (num_parts,), np.uint8)
    lib().oracle_spread_patches(_p(X, _i64), X.shape[0], X.shape[1], int(r0), int(r1), int(num_parts), _p(out, _u8))
    return out


spread_waliji_patches = spread_patches


def spread_graph_patches(X, part_graph):
    """code_parts.pyx:459-503 (in place)"""
    assert X.dtype == np.uint8 and X.flags.c_contiguous
    g = _c(part_graph, np.uint8)
    lib().oracle_spread_graph_patches(_p(X, _u8), X.shape[0], X.shape[1], X.shape[2], _p(g, _u8))


def get_parts_mask(X, M, S, threshold):
    """code_parts.pyx:297-348 (window-relative mask)"""
    X = _c(X, np.uint8); M = _c(M, np.uint8); S = _c(S, np.float64)
    T, F, E = X.shape
    ph, pw = M.shape
    counts = np.zeros((T - ph + 1, F - pw + 1), np.uint32)
    f = lib().oracle_get_parts_mask
    f.restype = ctypes.c_int
    null8, nullf, nullu = ctypes.POINTER(_u8)(), ctypes.POINTER(_f64)(), ctypes.POINTER(_u32)()
    n = f(_p(X, _u8), T, F, E, _p(M, _u8), ph, pw, _p(S, _f64), int(threshold), _p(counts, _u32), null8, nullf, nullu)
    patches = np.zeros((n, ph, pw, E), np.uint8)
    S_patches = np.zeros((n, ph, pw), np.float64)
    locs = np.zeros((n, 2), np.uint32)
    f(_p(X, _u8), T, F, E, _p(M, _u8), ph, pw, _p(S, _f64), int(threshold), _p(counts, _u32),
      _p(patches, _u8), _p(S_patches, _f64), _p(locs, _u32))
    return patches, S_patches, counts, locs


def get_max_above_threshold(detection_vals, detection_length, threshold, C0):
    """roc_functions.py:177-186 (the expression itself is valid Python 3 and is kept verbatim in meaning)"""
    l = detection_length - C0 + 1
    v = np.asarray(detection_vals)
    return np.arange(1, l - 1, dtype=np.uint16)[(v[1:l - 1] >= v[:l - 2]) * (v[1:l - 1] >= v[2:l]) * (v[1:l - 1] >= threshold)]


def cluster_times(times, C0, C1):
    """cluster_times.pyx:133-219: times uint16 (n,), strictly increasing -> partition uint8 (n+1,)"""
    times = _c(times, np.uint16)
    part = np.zeros(times.shape[0] + 1, np.uint8)
    lib().oracle_cluster_times(_p(times, ctypes.c_uint16), times.shape[0], ctypes.c_float(C0), ctypes.c_float(C1), _p(part, _u8))
    return part


def get_cluster_starts_ends(potential_detect_times, cluster_partition):
    """roc_functions.py:187-212: (start, end) = (first peak, last peak + 1) of every cluster"""
    out, t, d = [], 0, len(potential_detect_times)
    while t < d:
        start = potential_detect_times[t]
        t += 1
        while cluster_partition[t] != 2:
            t += 1
        out.append((start, potential_detect_times[t - 1] + 1))
    return tuple(out)


def load_ref_cluster_times():
    """compiled reference cluster_times (oracle/_ref) or None"""
    ref = os.path.join(HERE, "_ref")
    if not os.path.isdir(ref):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        m = importlib.import_module("cluster_times")
    except ImportError:
        return None
    return m if os.path.dirname(os.path.abspath(m.__file__)) == ref else None


# ------------------------------------------------------------------------ detection
def detect(F, LF):
    """compute_likelihood_linear_filter.pyx:135-176"""
    F = _c(F, np.uint8); LF = _c(LF, np.float32)
    out = np.empty(F.shape[0], np.float32)
    lib().oracle_detect(_p(F, _u8), F.shape[0], F.shape[1], F.shape[2], _p(LF, _f32), LF.shape[0], _p(out, _f32))
    return out


def detect_float(F, LF):
    """compute_likelihood_linear_filter.pyx:178-219; 2-D inputs give detect_float2d / detect_spectral"""
    F = _c(F, np.float32); LF = _c(LF, np.float32)
    if F.ndim == 2:
        F = F[:, :, None]; LF = LF[:, :, None]
    out = np.empty(F.shape[0], np.float32)
    lib().oracle_detect_float(_p(F, _f32), F.shape[0], F.shape[1], F.shape[2], _p(LF, _f32), LF.shape[0], _p(out, _f32))
    return out


detect_float2d = detect_float
detect_spectral = detect_float


def detect_float_max(F, LF, cs):
    """compute_likelihood_linear_filter.pyx:221-278"""
    F = _c(F, np.float32); LF = _c(LF, np.float32); cs = _c(cs, np.float32)
    out = np.empty(F.shape[0], np.float32)
    arg = np.empty(F.shape[0], np.uint16)
    lib().oracle_detect_float_max(_p(F, _f32), F.shape[0], F.shape[1], F.shape[2], _p(LF, _f32), LF.shape[0],
                                  LF.shape[1], _p(cs, _f32), _p(out, _f32), _p(arg, _u16))
    return out, arg


def detect_examples(F, LF, lengths, max_length):
    """compute_likelihood_linear_filter.pyx:79-131 (intended float64 semantics)"""
    F = _c(F, np.uint8); LF = _c(LF, np.float64); lengths = _c(lengths, np.int64)
    out = np.empty((F.shape[0], int(max_length)), np.float64)
    lib().oracle_detect_examples(_p(F, _u8), F.shape[0], F.shape[1], F.shape[2], F.shape[3], _p(LF, _f64),
                                 LF.shape[0], _p(lengths, _i64), int(max_length), _p(out, _f64), ctypes.POINTER(_f64)())
    return out


def get_max_detect_examples(F, LF, lengths):
    """compute_likelihood_linear_filter.pyx:21-76"""
    F = _c(F, np.uint8); LF = _c(LF, np.float64); lengths = _c(lengths, np.int64)
    mx = np.empty(F.shape[0], np.float64)
    lib().oracle_detect_examples(_p(F, _u8), F.shape[0], F.shape[1], F.shape[2], F.shape[3], _p(LF, _f64),
                                 LF.shape[0], _p(lengths, _i64), 0, ctypes.POINTER(_f64)(), _p(mx, _f64))
    return mx


# ------------------------------------------------------------------- edge spreading
def _spread_edges(T, direction, spread_length):
    """edge_signal_proc.py:1406-1417; returns a new array like the reference"""
    T = _c(T, np.float64)
    out = np.empty_like(T)
    lib().oracle_spread_edges(_p(T, _f64), T.shape[0], T.shape[1], int(direction[0]), int(direction[1]),
                              int(spread_length), _p(out, _f64))
    return out


def spread_edgemap(T, edge_feature_row_breaks, edge_orientation, spread_length=3):
    """edge_signal_proc.py:1397-1404 (in place on a float array of shape (rows, time))"""
    for b in range(1, len(edge_feature_row_breaks)):
        s, e = int(edge_feature_row_breaks[b - 1]), int(edge_feature_row_breaks[b])
        T[s:e, :] = _spread_edges(T[s:e], edge_orientation[b - 1], spread_length)


def spread_edgemap_all(T, edge_feature_row_breaks, spread_length):
    """'all' branch, edge_signal_proc.py:1623-1632: box maximum filter, constant-0 border"""
    for b in range(1, len(edge_feature_row_breaks)):
        s, e = int(edge_feature_row_breaks[b - 1]), int(edge_feature_row_breaks[b])
        blk = _c(T[s:e], np.float64)
        out = np.empty_like(blk)
        lib().oracle_spread_edges_box(_p(blk, _f64), blk.shape[0], blk.shape[1], int(spread_length), _p(out, _f64))
        T[s:e, :] = out


def spread_edge_blocks(X, spread_length, block_length, edge_orientations=None):
    """The spreading step of _edge_map_threshold_segments (edge_signal_proc.py:1617-1663) on an already binary edge
    tensor X uint8 (T, F, 8): every block_length-frame block is spread on its own; a ragged last block is the tail of
    the spread window of the final block_length frames (:1595-1633).  Returns a new (T, F, 8) uint8 array."""
    ori = EDGE_ORIENTATIONS if edge_orientations is None else np.asarray(edge_orientations).astype(int)
    X = np.asarray(X)
    T = X.shape[0]
    out = np.zeros_like(X)

    def one(win):
        res = np.empty_like(win)
        for e in range(win.shape[2]):
            res[:, :, e] = _spread_edges(win[:, :, e].T.astype(np.float64), ori[e], spread_length).T
        return res

    if block_length <= 0:
        return one(X)
    nfull, last = T // block_length, T % block_length
    for blk in range(nfull):
        out[blk * block_length:(blk + 1) * block_length] = one(X[blk * block_length:(blk + 1) * block_length])
    if last > 0:
        out[-last:] = one(X[-block_length:])[-last:]
    return out


# --------------------------------------------------------- edge front end (N4, N3)
ABST_THRESHOLD = np.array([0.025, 0.025, 0.015, 0.015, 0.02, 0.02, 0.02, 0.02])   # edge_signal_proc.py:1187
# orientation table the extractor returns (:1751-1821): (frequency step, time step) per block
EDGE_ORIENTATIONS = np.array([[0, 1], [0, -1], [1, 0], [-1, 0], [1, 1], [-1, -1], [-1, 1], [1, -1]])


def _edge_map_no_threshold(S):
    """edge_signal_proc.py:1697-1836 (time_axis_0=False, 2-D output).  For each of the four
    directions: a first-difference field D, kept where it is a (weak) local extremum of the right
    sign among its two neighbours along the direction; positive member then negative member."""
    S = np.asarray(S, np.float64)
    H, W = S.shape
    E = np.zeros((8 * H, W))
    # (D, rows of E written, columns of E written, neighbour before, neighbour after) as slices of D
    mid, lo, hi = slice(1, -1), slice(None, -2), slice(2, None)
    fields = (
        (S[:, 1:] - S[:, :-1], slice(0, H), slice(1, -2), (slice(None), lo), (slice(None), hi), (slice(None), mid)),    # :1744-1759
        (S[1:, :] - S[:-1, :], slice(1, H - 2), slice(None), (lo, slice(None)), (hi, slice(None)), (mid, slice(None))),  # :1764-1779
        (S[1:, 1:] - S[:-1, :-1], slice(1, H - 2), slice(1, -2), (lo, lo), (hi, hi), (mid, mid)),                        # :1784-1799
        (S[:-1, 1:] - S[1:, :-1], slice(1, H - 2), slice(1, -2), (hi, lo), (lo, hi), (mid, mid)),                        # :1804-1819
    )
    for pair, (D, rows, cols, before, after, centre) in enumerate(fields):
        C = D[centre]
        for sign, e in ((1.0, 2 * pair), (-1.0, 2 * pair + 1)):
            if sign > 0:
                if not np.any(D > 0):
                    continue
                keep = (C >= D[before]) * (C >= D[after]) * (C >= 0)
                val = C * keep
            else:
                if not np.any(D < 0):
                    continue
                keep = (C <= D[before]) * (C <= D[after]) * (C <= 0)
                val = -C * keep
            E[e * H + (rows.start or 0):e * H + (rows.stop if rows.stop is not None else H), cols] = val
    breaks = np.arange(9) * H
    return E, breaks, EDGE_ORIENTATIONS.copy()


def threshold_edge_block(E_block, quantile_level, abst_threshold):
    """edge_signal_proc.py:1261-1307 with convert_to_binary=True: returns (binary block, tau or None)."""
    pos = E_block > 0
    vals = np.sort(E_block[pos].ravel())
    if vals.shape[0] == 0:
        return np.zeros(E_block.shape), None
    tau = vals[int(quantile_level * vals.shape[0])]
    return (pos & (E_block >= max(tau, abst_threshold))).astype(np.float64), tau


def threshold_edgemap(E, quantile_level, edge_feature_row_breaks, abst_threshold=ABST_THRESHOLD):
    """edge_signal_proc.py:1184-1256 without critical bands: returns (binary map, taus (8,))."""
    out, taus = np.empty(E.shape), np.zeros(len(edge_feature_row_breaks) - 1)
    for e in range(len(edge_feature_row_breaks) - 1):
        a, b = int(edge_feature_row_breaks[e]), int(edge_feature_row_breaks[e + 1])
        out[a:b], tau = threshold_edge_block(E[a:b], quantile_level, abst_threshold[e])
        taus[e] = 0.0 if tau is None else tau
    return out, taus


def _edge_map_threshold_segments(E, block_length, spread_length, threshold=.7, edge_orientations=EDGE_ORIENTATIONS,
                                 edge_feature_row_breaks=None, abst_threshold=ABST_THRESHOLD, spread_type='unidirectional'):
    """edge_signal_proc.py:1537-1675, in place on E (raw magnitudes -> binary, spread per block).  The ragged
    last block is thresholded and spread on the window of the final block_length frames (:1595-1633)."""
    length = E.shape[1]
    if edge_feature_row_breaks is None:
        edge_feature_row_breaks = np.arange(9) * (E.shape[0] // 8)

    def one(window):
        binary, _ = threshold_edgemap(window, threshold, edge_feature_row_breaks, abst_threshold)
        if spread_length > 0 and spread_type == 'unidirectional':
            spread_edgemap(binary, edge_feature_row_breaks, edge_orientations, spread_length)
        elif spread_length > 0 and spread_type == 'all':
            spread_edgemap_all(binary, edge_feature_row_breaks, spread_length)
        return binary

    nfull, last = length // block_length, length % block_length
    tail = one(E[:, -block_length:].copy())[:, -last:] if last > 0 else None
    for blk in range(nfull):
        E[:, blk * block_length:(blk + 1) * block_length] = one(E[:, blk * block_length:(blk + 1) * block_length])
    if last > 0:
        E[:, -last:] = tail


# ------------------------------------------------------------- linear filter / LLR
def construct_linear_filter(T, bgd, min_prob=.01, all_cs=False):
    """estimate_template.py:230-251.  T (L, ...) probabilities, bgd (...) one frame."""
    T = np.clip(T, min_prob, 1 - min_prob)
    Bgd = np.tile(bgd, (T.shape[0],) + tuple(1 for _ in range(T.ndim - 1)))
    T_inv = 1. - T
    Bgd_inv = 1. - Bgd
    C_exp_inv = T_inv / Bgd_inv
    if all_cs:
        c = np.cumsum(np.log(C_exp_inv.reshape(len(C_exp_inv), -1)).sum(1))
    else:
        c = np.log(C_exp_inv).sum()
    expW = (T / Bgd) / C_exp_inv
    return np.log(expW).astype(np.float32), c


def construct_linear_filters(Ts, bgd, all_cs=False, min_prob=.01):
    """estimate_template.py:156-172 (Bernoulli branch)"""
    return tuple(construct_linear_filter(np.clip(T, min_prob, 1 - min_prob), np.clip(bgd, min_prob, 1 - min_prob),
                                         all_cs=all_cs) for T in Ts)


def score_template_background_section(template, bgd, E, front_bgd_pad=0, back_bgd_pad=0):
    """test_template.py:25-41.  template (H, L), bgd (H,), E (H, L)."""
    assert E.shape[1] == template.shape[1]
    T = np.hstack((np.tile(bgd, (front_bgd_pad, 1)).transpose(), template,
                   np.tile(bgd, (back_bgd_pad, 1)).transpose()))
    L = T.shape[1]
    U_bgd = np.tile(bgd, (L, 1)).transpose()
    C_exp_inv_long = (1. - T) / (1. - U_bgd)
    C = np.log(C_exp_inv_long).sum()
    expW = (T / U_bgd) / C_exp_inv_long
    return (E * np.log(expW)).sum(), C


# ------------------------------------------------------------------ Bernoulli mixture
class BernoulliMixture:
    """Restatement of bernoulli_mixture.py:9-311 (prints dropped, xrange->range, // for
    Python-2 integer division).  dtypes follow the reference line by line."""

    def __init__(self, num_mix, data_mat, init_type='unif_rand', init_seed=0, init_centers=None,
                 min_probability=.01):
        self.num_mix = num_mix                                            # :76
        self.num_data = data_mat.shape[0]
        self.data_shape = data_mat.shape[1:]
        self.data_length = int(np.prod(data_mat.shape[1:]))
        self.data_mat = data_mat.reshape(self.num_data, self.data_length)  # :82
        self.iterations = 0
        self.seed = init_seed
        self.min_probability = np.float32(min_probability)                # :89
        self.not_data_mat = 1 - self.data_mat                             # :92
        self.weights = 1. / num_mix * np.ones(num_mix, dtype=np.float32)  # :95
        self.init_affinities_templates(init_type, init_centers=init_centers)

    def run_EM(self, tol, min_probability=0.05, hard_assignment=False, max_iter=500, min_iter=10):
        self.hard_assignment = hard_assignment                            # :131
        self.min_probability = np.float32(min_probability)
        new_loglikelihood = self._compute_loglikelihoods()                # :135
        loglikelihood = new_loglikelihood / (1 - tol) - 10                # :139
        self.iterations = 0
        self.loglikelihood_trace = [new_loglikelihood]
        while (np.abs((loglikelihood - new_loglikelihood) / loglikelihood) > tol
               and self.iterations < max_iter) or self.iterations < min_iter:   # :144
            loglikelihood = new_loglikelihood
            self.M_step()
            new_loglikelihood = self._compute_loglikelihoods()
            self.loglikelihood_trace.append(new_loglikelihood)
            self.iterations += 1
        self.loglikelihood = new_loglikelihood
        self.set_templates()                                              # :160

    def M_step(self):                                                     # :177-184
        self.weights = np.mean(self.affinities, axis=0)
        self.work_templates = np.dot(self.affinities.T, self.data_mat).astype(np.float32)
        self.work_templates /= self.num_data
        self.work_templates /= self.weights.reshape((self.num_mix, 1))
        self.threshold_templates()
        self.log_templates = np.log(self.work_templates).astype(np.float32)
        self.log_invtemplates = np.log(1 - self.work_templates).astype(np.float32)

    def threshold_templates(self):                                        # :192-193
        self.work_templates = np.clip(self.work_templates, self.min_probability, 1 - self.min_probability)

    def init_affinities_templates(self, init_type, init_centers=None):    # :195-237
        if init_type == 'unif_rand':
            self.affinities = np.zeros((self.num_data, self.num_mix))
            self.work_templates = np.zeros((self.num_mix, self.data_length))
            for mix_id in range(self.num_mix):
                self.affinities[self.num_mix * np.arange(self.num_data // self.num_mix) + mix_id, mix_id] = 1.
                self.work_templates[mix_id] = np.mean(self.data_mat[self.affinities[:, mix_id] == 1], axis=0)
                self.threshold_templates()
        elif init_type == 'preset' and init_centers is not None:
            self.num_mix = init_centers.shape[0]
            self.work_templates = np.clip(init_centers.reshape(self.num_mix, self.data_length),
                                          self.min_probability, 1 - self.min_probability)
            self.affinities = np.zeros((self.num_data, self.num_mix))
        elif init_type == 'specific':
            self.affinities = np.zeros((self.num_data, self.num_mix))
            self.work_templates = np.zeros((self.num_mix, self.data_length))
            for mix_id in range(self.num_mix):
                self.affinities[self.num_mix * np.arange(self.num_data // self.num_mix)[1] + mix_id, mix_id] = 1.
                self.work_templates[mix_id] = np.mean(self.data_mat[self.affinities[:, mix_id] == 1], axis=0)
                self.threshold_templates()
        self.work_templates = np.clip(self.work_templates, self.min_probability, 1 - self.min_probability)
        self.log_templates = np.log(self.work_templates)
        self.log_invtemplates = np.log(1 - self.work_templates)

    def set_templates(self):                                              # :245-248
        self.templates = self.work_templates.reshape((self.num_mix,) + self.data_shape)
        self.log_templates = np.log(self.templates)
        self.log_invtemplates = np.log(1 - self.templates)

    def _compute_loglikelihoods(self):                                    # :263-280
        template_logscores = self.get_template_loglikelihoods()
        loglikelihoods = template_logscores + np.tile(np.log(self.weights), (self.num_data, 1))
        max_vals = np.amax(loglikelihoods, axis=1)
        logmarginals_adj = np.sum(np.exp(loglikelihoods - np.tile(max_vals, (self.num_mix, 1)).transpose()), axis=1)
        loglikelihood = np.sum(np.log(logmarginals_adj)) + np.sum(max_vals)
        if self.hard_assignment:
            self.affinities[:] = 0
            self.affinities[np.arange(self.num_data), np.argmax(template_logscores, 1)] = 1
            return loglikelihood
        self.affinities = np.exp(loglikelihoods - np.tile(logmarginals_adj + max_vals,
                                                          (self.num_mix, 1)).transpose()).astype(np.float32)
        self.affinities /= np.tile(np.sum(self.affinities, axis=1), (self.num_mix, 1)).transpose().astype(np.float32)
        return loglikelihood

    def get_template_loglikelihoods(self):                                # :287-291
        return np.dot(self.data_mat, self.log_templates.T) + np.dot(self.not_data_mat, self.log_invtemplates.T)


class Bernoulli_Mixture:
    """Restatement of bernoulli_em.py:8-131 (first-version EM; float64 throughout, clip
    fixed at [.05,.95], absolute-difference stopping rule)."""

    def __init__(self, num_mix, data_mat, init_type='unif_rand', opt_type='expected'):
        self.num_mix = num_mix
        self.height, self.width = data_mat.shape[1:]
        self.data_mat = data_mat
        self.num_data = data_mat.shape[0]
        self.weights = 1. / num_mix * np.ones(num_mix)
        self.init_affinities_templates(init_type)

    def run_EM(self, tol):                                                # :22-34
        loglikelihood = -np.inf
        new_loglikelihood = self.compute_loglikelihoods()
        self.iterations = 0
        while new_loglikelihood - loglikelihood > tol:
            loglikelihood = new_loglikelihood
            self.M_step()
            new_loglikelihood = self.compute_loglikelihoods()
            self.iterations += 1
        self.loglikelihood = new_loglikelihood

    def M_step(self):                                                     # :36-45
        self.weights = np.mean(self.affinities, axis=0)
        for mix_id in range(self.num_mix):
            self.templates[mix_id] = np.sum(self.data_mat * self.affinities[:, mix_id][:, None, None], axis=0)
            self.templates[mix_id] /= (self.weights[mix_id] * self.num_data)
        self.threshold_templates()
        self.log_templates = np.log(self.templates)
        self.log_invtemplates = np.log(1 - self.templates)

    def threshold_templates(self):                                        # :48-49
        self.templates = np.maximum(np.minimum(self.templates, .95), .05)

    def init_affinities_templates(self, init_type):                       # :51-79
        self.affinities = np.zeros((self.num_data, self.num_mix))
        self.templates = np.zeros((self.num_mix, self.height, self.width))
        for mix_id in range(self.num_mix):
            if init_type == 'unif_rand':
                idx = self.num_mix * np.arange(self.num_data // self.num_mix) + mix_id
            else:  # 'specific'
                idx = self.num_mix * np.arange(self.num_data // self.num_mix)[1] + mix_id
            self.affinities[idx, mix_id] = 1.
            self.templates[mix_id] = np.mean(self.data_mat[self.affinities[:, mix_id] == 1], axis=0)
            self.threshold_templates()
        self.log_templates = np.log(self.templates)
        self.log_invtemplates = np.log(1 - self.templates)

    def compute_loglikelihoods(self):                                     # :103-114
        template_logscores = self.get_template_loglikelihoods(self.data_mat)
        loglikelihoods = template_logscores + np.tile(np.log(self.weights), (self.num_data, 1))
        max_vals = np.amax(loglikelihoods, axis=1)
        logmarginals_adj = np.sum(np.exp(loglikelihoods - np.tile(max_vals, (self.num_mix, 1)).transpose()), axis=1)
        loglikelihood = np.sum(np.log(logmarginals_adj)) + np.sum(max_vals)
        self.affinities = np.exp(loglikelihoods - np.tile(logmarginals_adj + max_vals, (self.num_mix, 1)).transpose())
        self.affinities /= np.tile(np.sum(self.affinities, axis=1), (self.num_mix, 1)).transpose()
        return loglikelihood

    def get_template_loglikelihood(self, datum):                          # :116-120
        return np.sum(self.log_templates[:, datum > .5], axis=1) + np.sum(self.log_invtemplates[:, datum < .5], axis=1)

    def get_template_loglikelihoods(self, data_mat):                      # :122-127
        return np.array([self.get_template_loglikelihood(x) for x in data_mat])
