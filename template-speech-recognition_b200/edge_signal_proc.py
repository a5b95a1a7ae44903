"""Drop-in for the edge functions of ``edge_signal_proc`` on the scoring path: edge spreading
(template_speech_rec/edge_signal_proc.py:1397-1417 and the 'all' branch :1623-1632) and the edge
front end before it -- 8-orientation edge extraction ``_edge_map_no_threshold`` (:1697-1836) and the
per-block quantile thresholding ``threshold_edge_block`` / ``threshold_edgemap`` /
``_edge_map_threshold_segments`` (:1184-1307, :1537-1675).

The reference keeps the 8 orientation blocks stacked along the rows of a float (rows, time)
array.  These wrappers pack the blocks into one byte per (time, frequency) location, run the
bit-packed max-filter kernel and write 0/1 back.  The map must be binary (the reference
spreads thresholded maps, edge_signal_proc.py:1650-1660); other values raise ValueError.
"""
import numpy as np

from . import ops
from ._np import to_dev


def _as_binary(T):
    T = np.asarray(T)
    if T.ndim != 2:
        raise ValueError("edge map must be 2-D (feature rows, time)")
    if not np.all((T == 0) | (T == 1)):
        raise ValueError("edge map must be binary (0/1) for the bit-packed spreading kernel")
    return T


def _blocks(T, breaks):
    breaks = [int(b) for b in np.asarray(breaks).ravel()]
    nb = len(breaks) - 1
    if nb < 1 or nb > 8:
        raise ValueError("need between 1 and 8 orientation blocks, got %d" % nb)
    hs = [breaks[i + 1] - breaks[i] for i in range(nb)]
    if min(hs) < 0 or breaks[0] < 0 or breaks[-1] > T.shape[0]:
        raise ValueError("edge_feature_row_breaks out of range")
    return breaks, nb, hs


def _run(T, breaks, dirs, radius, mode):
    breaks, nb, hs = _blocks(T, breaks)
    H, W = max(hs), T.shape[1]
    if H == 0 or W == 0:
        return
    X = np.zeros((1, W, H, 8), np.uint8)                     # (time, frequency, orientation)
    for e in range(nb):
        X[0, :, :hs[e], e] = T[breaks[e]:breaks[e + 1]].T != 0
    bits = ops.pack_edges(to_dev(X))
    out = ops.unpack_edges(ops.spread_edges(bits, radius, orientations=dirs, mode=mode), 8)[0].cpu().numpy()
    for e in range(nb):
        T[breaks[e]:breaks[e + 1], :] = out[:, :hs[e], e].T


def spread_edgemap(T, edge_feature_row_breaks, edge_orientation, spread_length=3):
    """:1397-1404, in place.  edge_orientation[e] = (row step, time step) of block e."""
    _as_binary(T)
    ori = np.asarray(edge_orientation)
    breaks, nb, _ = _blocks(T, edge_feature_row_breaks)
    if ori.ndim != 2 or ori.shape[0] < nb or ori.shape[1] != 2 or np.any(ori != np.round(ori)):
        raise ValueError("edge_orientation must hold one integer (row step, time step) pair per block")
    dirs = np.zeros((8, 2), np.int8)
    dirs[:nb] = ori[:nb].astype(np.int8)
    if spread_length > 0:
        _run(T, edge_feature_row_breaks, dirs, int(spread_length), "line")


def _spread_edges(T, direction, spread_length):
    """:1406-1417: one orientation block, returns a new array."""
    T = _as_binary(T)
    out = np.array(T, dtype=np.float64)
    dirs = np.zeros((8, 2), np.int8)
    dirs[0] = (int(direction[0]), int(direction[1]))
    if spread_length > 0:
        _run(out, [0, T.shape[0]], dirs, int(spread_length), "line")
    return out


def spread_edgemap_all(T, edge_feature_row_breaks, spread_length):
    """The 'all' branch of _edge_map_threshold_segments (:1623-1632): (2s+1)^2 box maximum filter
    per orientation block with a constant-0 border, in place."""
    _as_binary(T)
    if spread_length > 0:
        _run(T, edge_feature_row_breaks, None, int(spread_length), "box")


# ------------------------------------------------------------------ edge front end (N4, N3)
import torch  # noqa: E402  (device memory only)

_ABST = np.array([0.025, 0.025, 0.015, 0.015, 0.02, 0.02, 0.02, 0.02])
_SEGMENT_ORIENTATIONS = np.array([[1., 0.], [-1., 0.], [0., 1.], [0., -1.], [1., 1.], [-1., -1.], [1., -1.], [-1., 1.]])


def compute_edge_feature_row_breaks(base_height, edge_orientations):
    """:1679-1693 -- blocks that take no vertical difference are three rows taller."""
    heights = np.zeros(9)
    heights[1:] = base_height + 3 * (np.asarray(edge_orientations)[:, 0] == 0)
    return np.cumsum(heights)


def _edge_map_no_threshold(S, return_3_dimensions=False, time_axis_0=False):
    """:1697-1836.  Returns (E, edge_feature_row_breaks, edge_orientations); E stacks the 8 orientation
    blocks along the rows of a (8*H, time) float64 array (the axis juggling of the reference's
    time_axis_0 / return_3_dimensions options is reproduced)."""
    S = np.asarray(S, dtype=np.float64)
    if S.ndim != 2:
        raise ValueError("S must be 2-D")
    if time_axis_0:
        S = S.T
    H, W = S.shape
    E = ops.edge_map(to_dev(S[None]))[0].cpu().numpy()                     # (8, H, W)
    breaks = np.arange(9, dtype=np.float64) * H
    ori = ops.DEFAULT_EDGE_ORIENTATIONS.astype(np.float64)
    if time_axis_0:
        # :1822-1834: E.T twice for the 2-D case; (8,T,H).T for the 3-D case
        E = np.ascontiguousarray(E.transpose(1, 2, 0)) if return_3_dimensions else E.reshape(8 * H, W)
    elif not return_3_dimensions:
        E = E.reshape(8 * H, W)
    return E, breaks, ori


def _segments_on_device(E, breaks, block_length, quantile, abst_threshold):
    """rows of E grouped by `breaks` -> (1,8,Hmax,T) magnitudes on the device (shorter blocks zero padded:
    zeros are never edges and never enter a quantile)."""
    breaks, nb, hs = _blocks(E, breaks)
    if nb != 8:
        raise ValueError("the quantile thresholding needs 8 orientation blocks, got %d" % nb)
    H, W = max(hs), E.shape[1]
    mags = np.zeros((1, 8, H, W))
    for e in range(8):
        mags[0, e, :hs[e]] = E[breaks[e]:breaks[e + 1]]
    return breaks, hs, ops.threshold_edges(to_dev(mags), block_length, quantile, abst_threshold, want_taus=True)


def _write_back(E, breaks, hs, bits, convert_to_binary):
    out = ops.unpack_edges(bits, 8)[0].cpu().numpy()                       # (T, Hmax, 8)
    for e in range(8):
        binary = out[:, :hs[e], e].T
        blk = E[breaks[e]:breaks[e + 1]]
        blk[...] = binary if convert_to_binary else np.where(binary != 0, blk, 0.)


def threshold_edge_block(E_block, quantile_level, report_level, abst_threshold, convert_to_binary=True):
    """:1261-1307, in place: keep (as 1, or as the magnitude) the positive entries >= max(tau, abst_threshold),
    tau = sorted(positive entries)[int(quantile_level * n)]."""
    if E_block.ndim != 2:
        raise ValueError("E_block must be 2-D")
    if not np.any(E_block > 0):
        E_block[:] = 0
        return E_block                                                      # as the reference, even with report_level
    mags = np.zeros((1, 8) + E_block.shape)
    mags[0, 0] = E_block
    abst = np.zeros(8)
    abst[0] = abst_threshold
    bits, _, taus = ops.threshold_edges(to_dev(mags), E_block.shape[1], quantile_level, abst, want_taus=True)
    binary = (bits[0].cpu().numpy().T & 1)
    E_block[...] = binary if convert_to_binary else np.where(binary != 0, E_block, 0.)
    return (E_block, float(taus[0, 0, 0])) if report_level else E_block


def threshold_edgemap(E, quantile_level, edge_feature_row_breaks, report_level=False, abst_threshold=_ABST,
                      critical_bands=None, convert_to_binary=True):
    """:1184-1256: every orientation block of E thresholded as one time block, in place (the reference
    writes through views of E); returns (E, thresholds (8,)) when report_level, else None."""
    if critical_bands is not None:
        raise NotImplementedError("critical_bands: the reference slices E[start+cb[0]:end+cb[1]] (:1231), which is "
                                  "not a partition of the block; not reproduced")
    if E.shape[1] == 0:
        return (E, np.zeros(8)) if report_level else None
    breaks, hs, (bits, _, taus) = _segments_on_device(E, edge_feature_row_breaks, E.shape[1], quantile_level, abst_threshold)
    _write_back(E, breaks, hs, bits, convert_to_binary)
    if report_level:
        return E, taus[0, 0].cpu().numpy()


def _edge_map_threshold_segments(E, block_length, spread_length, threshold=.7, edge_orientations=_SEGMENT_ORIENTATIONS,
                                 edge_feature_row_breaks=None, abst_threshold=_ABST, verbose=False, critical_bands=None,
                                 return_thresholds=False, spread_type='unidirectional', convert_to_binary=True):
    """:1537-1675, in place: per `block_length`-frame block, quantile-threshold each orientation block and
    spread it (never across a block boundary); the ragged last block is recomputed on the window of the
    final block_length frames (:1595-1633).  With return_thresholds the reference returns after the first
    full block (:1674-1675 sits inside the loop); here every block is processed and the list holds the
    ragged window first, then the full blocks in order."""
    if critical_bands is not None:
        raise NotImplementedError("critical_bands (see threshold_edgemap)")
    if spread_type not in ('unidirectional', 'all'):
        raise ValueError("spread_type must be 'unidirectional' or 'all'")
    height, length = E.shape
    if edge_feature_row_breaks is None:
        base_height = (height - 6) // 8
        assert base_height * 8 == height - 6
        edge_feature_row_breaks = compute_edge_feature_row_breaks(base_height, edge_orientations)
    block_length = int(block_length)
    if length == 0:
        return [] if return_thresholds else None
    breaks, hs, (bits, tail, taus) = _segments_on_device(E, edge_feature_row_breaks, block_length, threshold, abst_threshold)
    if spread_length > 0:
        if not convert_to_binary:
            raise NotImplementedError("spreading real-valued (convert_to_binary=False) maps is not on the scoring path")
        ori = np.asarray(edge_orientations)
        if spread_type == 'unidirectional':
            if ori.shape != (8, 2) or np.any(ori != np.round(ori)):
                raise ValueError("edge_orientations must hold 8 integer (row step, time step) pairs")
            bits = ops.spread_edge_blocks(bits, tail, int(spread_length), block_length, ori.astype(np.int8), "line")
        else:
            bits = ops.spread_edge_blocks(bits, tail, int(spread_length), block_length, None, "box")
    _write_back(E, breaks, hs, bits, convert_to_binary)
    if verbose or return_thresholds:
        taus = taus[0].cpu().numpy()
        order = ([len(taus) - 1] if length % block_length else []) + list(range(length // block_length))
        views = [taus[i] for i in order]
        if verbose:
            for v in views:
                print(v)
        if return_thresholds:
            return views
