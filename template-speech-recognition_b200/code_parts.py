"""Drop-in for the reference Cython module ``code_parts`` (template_speech_rec/code_parts.pyx).

Same function names, argument order, dtypes and output shapes; NumPy arrays in, freshly
allocated NumPy arrays out.  Every function stages its inputs to the GPU, runs the
sm_100a kernels of libtsr_b200.so and copies the result back -- there is no CPU path.
For throughput use the device-resident API in ``ops`` / ``pipeline`` instead; these
wrappers exist so the reference's callers (get_train_data.py:183-189, 257-266, 2600-2628;
parts.py:160-163) run unchanged.
"""
import numpy as np
import torch

from . import ops
from ._np import Cache, device, digest, to_dev, typed

_parts_cache = Cache(4)
_INT_MAX = 2 ** 31 - 1


def _parts(log_parts, log_invparts, mask=None):
    """Validated, cached device copy of a part dictionary (arguments are checked before any GPU work)."""
    lp = typed(log_parts, np.float64, 4, "DTYPE_t")
    lip = typed(log_invparts, np.float64, 4, "DTYPE_t")
    if lip.shape != lp.shape:
        raise ValueError("log_parts and log_invparts must have the same shape")
    if lp.shape[3] > 8:
        raise ValueError("more than 8 edge types per location are not supported by the packed representation")
    key = digest(lp, lip) + (b"" if mask is None else digest(mask))
    return _parts_cache.get(key, lambda: ops.PartsModel(lp, lip, mask=mask, device=device()))


def _unsigned_threshold(t):
    # the reference compares `unsigned int count >= int threshold` in C: a negative threshold
    # converts to a huge unsigned value and never passes (code_parts.pyx:120)
    t = int(t)
    return _INT_MAX if t < 0 else min(t, _INT_MAX)


def _run(X, parts, threshold, mask_threshold=-1):
    X = typed(X, np.uint8, 3, "UINT_t")
    if X.shape[2] != parts.E:
        raise ValueError("X has %d edge types but the parts have %d" % (X.shape[2], parts.E))
    if X.shape[0] < parts.ph or X.shape[1] < parts.pw:
        raise ValueError("X (%d,%d) is smaller than the part window (%d,%d)" % (X.shape[0], X.shape[1], parts.ph, parts.pw))
    bits = ops.pack_edges(to_dev(X[None]))
    codes, scores, counts = ops.code_parts(bits, parts, threshold, mask_threshold, want_codes=True,
                                           want_scores=True, want_counts=True)
    gate = (codes[0] != 0).cpu().numpy()
    return gate, scores[0].cpu().numpy().astype(np.float64), counts[0].cpu().numpy()


def code_parts(X, log_parts, log_invparts, threshold):
    """code_parts.pyx:70-135.  Returns float64 (T-ph+1, F-pw+1, K+1): slot 0 is -inf and slots
    1..K the part log-likelihoods where the window holds >= threshold edges, all zeros otherwise."""
    X = typed(X, np.uint8, 3, "UINT_t")
    parts = _parts(log_parts, log_invparts)
    gate, scores, _ = _run(X, parts, _unsigned_threshold(threshold))
    out = np.zeros(scores.shape[:2] + (parts.K + 1,), np.float64)
    out[..., 1:] = np.where(gate[..., None], scores, 0.0)
    out[..., 0] = np.where(gate, -np.inf, 0.0)
    return out


def code_parts_fast(X, log_parts, log_invparts, threshold):
    """code_parts.pyx:247-295: as code_parts but slot 0 = sum_k(score_k - 1) above threshold
    (:290-291) and 1.0 below (:293)."""
    X = typed(X, np.uint8, 3, "UINT_t")
    parts = _parts(log_parts, log_invparts)
    gate, scores, _ = _run(X, parts, _unsigned_threshold(threshold))
    out = np.zeros(scores.shape[:2] + (parts.K + 1,), np.float64)
    out[..., 1:] = np.where(gate[..., None], scores, 0.0)
    out[..., 0] = np.where(gate, (scores - 1.0).sum(-1), 1.0)
    return out


def code_parts_mask(X, M, log_parts, log_invparts, mask_threshold, absolute_threshold):
    """code_parts.pyx:353-404: gate = count >= absolute_threshold and mask_count > mask_threshold,
    mask_count counting edges under the 2-D window mask M (window-relative; the reference indexes
    M with absolute coordinates, which reads out of bounds -- see DESIGN.md)."""
    X = typed(X, np.uint8, 3, "UINT_t")
    M = typed(M, np.uint8, 2, "UINT_t")
    parts = _parts(log_parts, log_invparts, mask=M)
    mt = int(mask_threshold)
    if mt < 0:           # unsigned comparison in the reference: nothing is > a huge unsigned value
        mt = _INT_MAX
    gate, scores, _ = _run(X, parts, _unsigned_threshold(absolute_threshold), mt)
    out = np.zeros(scores.shape[:2] + (parts.K + 1,), np.float64)
    out[..., 1:] = np.where(gate[..., None], scores, 0.0)
    out[..., 0] = np.where(gate, (scores - 1.0).sum(-1), 1.0)
    return out


def code_parts_mag_aux(X_edges, X_mags, X_aux, log_parts, log_invparts, log_mag_parts, log_mag_invparts,
                       log_aux_parts, log_aux_invparts, X_thresh_mask):
    """code_parts.pyx:138-242: edge log-likelihoods plus band-magnitude and auxiliary Bernoulli terms,
    gated by X_thresh_mask at the window centre (:191-192; no edge-count threshold).  The edge
    contraction runs on the tensor kernel, the extra terms in tsr_code_parts_mag_aux_terms."""
    X = typed(X_edges, np.uint8, 3, "UINT_t")
    Xm = typed(X_mags, np.uint8, 2, "UINT_t")
    Xa = typed(X_aux, np.uint8, 2, "UINT_t")
    parts = _parts(log_parts, log_invparts)
    tabs = [typed(t, np.float64, 2, "DTYPE_t") for t in (log_mag_parts, log_mag_invparts, log_aux_parts, log_aux_invparts)]
    Tm = typed(X_thresh_mask, np.uint8, 2, "UINT_t")
    if X.shape[2] != parts.E:
        raise ValueError("X has %d edge types but the parts have %d" % (X.shape[2], parts.E))
    if Xm.shape[0] != X.shape[0] or Xa.shape[0] != X.shape[0] or Tm.shape != X.shape[:2]:
        raise ValueError("X_mags, X_aux and X_thresh_mask must follow the first dimensions of X_edges")
    bits = ops.pack_edges(to_dev(X[None]))
    _, scores, _ = ops.code_parts(bits, parts, 0, want_codes=False, want_scores=True)
    gate = ops.code_parts_mag_aux_terms(scores, to_dev(Xm[None]), to_dev(Xa[None]), to_dev(Tm[None]),
                                        *[to_dev(t) for t in tabs], parts.ph, parts.pw)
    gate = gate[0].cpu().numpy() != 0
    out = np.zeros(gate.shape + (parts.K + 1,), np.float64)
    out[..., 1:] = scores[0].cpu().numpy()
    out[..., 0] = np.where(gate, -np.inf, 0.0)
    return out


def get_parts_mask(X, M, S, threshold):
    """code_parts.pyx:297-348: every ph x pw x E patch whose masked edge count is >= threshold, in
    row-major order, with the matching spectrogram patches, the count map and the patch origins.
    Returns (patch_set uint8 (n,ph,pw,E), S_patch_set float64 (n,ph,pw), patch_counts uint32 (To,Fo),
    patch_locs uint32 (n,2)).  The 2-D mask is window-relative (see code_parts_mask)."""
    X = typed(X, np.uint8, 3, "UINT_t")
    M = typed(M, np.uint8, 2, "UINT_t")
    S = typed(S, np.float64, 2, "DTYPE_t")
    ph, pw = M.shape
    if X.shape[0] < ph or X.shape[1] < pw:
        raise ValueError("X (%d,%d) is smaller than the mask window (%d,%d)" % (X.shape[0], X.shape[1], ph, pw))
    if S.shape != X.shape[:2]:
        raise ValueError("S must have the first two dimensions of X")
    if X.shape[2] > 8:
        raise ValueError("more than 8 edge types per location are not supported by the packed representation")
    bits = ops.pack_edges(to_dev(X[None]))
    r = ops.extract_patches(bits, ph, pw, _unsigned_threshold(threshold), mask=to_dev((M != 0).astype(np.uint8)),
                            S=to_dev(S[None]), E=X.shape[2], want_bytes=True, want_bits=False)
    # the reference copies the raw bytes of X; nonzero values other than 1 are restored from X here
    locs = r["locs"].cpu().numpy().astype(np.uint32)[:, 1:]
    patches = r["patches"].cpu().numpy()
    if X.max(initial=0) > 1 and len(locs):
        ii = locs[:, 0, None, None] + np.arange(ph)[None, :, None]
        jj = locs[:, 1, None, None] + np.arange(pw)[None, None, :]
        patches = X[ii, jj]
    return patches, r["S_patches"].cpu().numpy(), r["counts"][0].cpu().numpy().astype(np.uint32), locs


def _check_codes(X, num_parts):
    X = typed(X, np.int64, 2, "long")
    if X.shape[0] > 65535 or X.shape[1] > 65535:
        raise ValueError("dimension > 65535: the reference holds sizes in uint16 and silently wraps "
                         "(code_parts.pyx:440-441); split the input")
    num_parts = int(num_parts)
    if not 1 <= num_parts <= 255:
        raise ValueError("num_parts must be in [1,255] for the uint8 code representation")
    if X.size and (X.min() < 0 or X.max() > num_parts):
        raise ValueError("part codes must lie in [0, num_parts]")
    return X, num_parts


def spread_patches(X, spread_0_dim, spread_1_dim, num_parts):
    """code_parts.pyx:406-457.  X int64 (D0,D1) codes (0 = background) -> uint8 (D0,D1,num_parts)."""
    X, K = _check_codes(X, num_parts)
    D0, D1 = X.shape
    if X.size == 0:
        return np.zeros((D0, D1, K), np.uint8)
    if spread_0_dim < 0 or spread_1_dim < 0:
        return np.zeros((D0, D1, K), np.uint8)   # empty window ranges in the reference loops
    codes = to_dev(X.astype(np.uint8)[None])
    fb = ops.spread_codes(codes, K, int(spread_0_dim), int(spread_1_dim))
    return ops.unpack_features(fb, D1 * K)[0].cpu().numpy().reshape(D0, D1, K)


def spread_graph_patches(X, part_graph):
    """code_parts.pyx:459-503: in place on X uint8 (D0,D1,K); returns None like the reference."""
    if not (isinstance(X, np.ndarray) and X.dtype == np.uint8 and X.ndim == 3):
        typed(X, np.uint8, 3, "uint8_t")
    G = typed(part_graph, np.uint8, 2, "uint8_t")
    D0, D1, K = X.shape
    if G.shape != (K, K):
        raise ValueError("part_graph must be (%d,%d)" % (K, K))
    if X.size == 0:
        return None
    fb = ops.pack_features(to_dev(np.ascontiguousarray(X).reshape(1, D0, D1 * K)))
    ops.spread_graph(fb, to_dev((G != 0).astype(np.uint8)), D1, K)
    res = ops.unpack_features(fb, D1 * K)[0].cpu().numpy().reshape(D0, D1, K)
    # the reference only ever writes ones (X[i,j,p2] = 1) and leaves other values untouched
    np.copyto(X, np.where(res != 0, np.where(X != 0, X, 1), X).astype(np.uint8))
    return None
