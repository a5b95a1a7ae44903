"""Drop-in for the edge-spreading functions of ``edge_signal_proc``
(template_speech_rec/edge_signal_proc.py:1397-1417 and the 'all' branch :1623-1632).

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
