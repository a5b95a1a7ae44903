"""Drop-in for the peak picking and peak clustering of the reference's ``roc_functions`` module
(template_speech_rec/roc_functions.py): the steps right after detection.  NumPy in, NumPy out;
the work runs on the GPU (tsr_detect_peaks + ordered compaction), there is no CPU path."""
import numpy as np

from . import ops
from ._np import to_dev


def get_max_above_threshold(detection_vals, detection_length, threshold, C0):
    """roc_functions.py:177-186: indices t in [1, l-2], l = detection_length - C0 + 1, where
    detection_vals[t] is >= both neighbours and >= threshold, as uint16 (the reference's dtype)."""
    v = np.ascontiguousarray(detection_vals, dtype=np.float32)
    if v.ndim != 1:
        raise ValueError("detection_vals must be one-dimensional")
    l = int(detection_length) - int(C0) + 1
    if l > v.shape[0]:
        raise ValueError("operands could not be broadcast together: detection_length - C0 + 1 exceeds len(detection_vals)")
    if v.shape[0] > 65535:
        raise ValueError("dimension > 65535: the reference returns uint16 indices; split the track")
    if l < 3:
        return np.zeros(0, np.uint16)
    lengths = to_dev(np.array([int(detection_length)], np.int32))
    bt, _ = ops.detect_peaks(to_dev(v[None]), float(threshold), int(C0), lengths)
    return bt[:, 1].cpu().numpy().astype(np.uint16)


def get_cluster_starts_ends(potential_detect_times, cluster_partition):
    """roc_functions.py:187-212 (host bookkeeping over a handful of peaks): (first peak, last peak + 1) per cluster."""
    out, t, d = [], 0, len(potential_detect_times)
    while t < d:
        start = potential_detect_times[t]
        t += 1
        while cluster_partition[t] != 2:
            t += 1
        out.append((start, potential_detect_times[t - 1] + 1))
    return tuple(out)


def perform_detection_time_clustering(potential_detect_times, C0, C1):
    """roc_functions.py:213-222: uint16 strictly increasing peak frames -> tuple of (start, end) clusters."""
    times = np.ascontiguousarray(potential_detect_times)
    if times.dtype != np.uint16 or times.ndim != 1:
        raise ValueError("Buffer dtype mismatch, expected 'npy_uint16' but got '%s'" % times.dtype)
    bt = np.zeros((times.shape[0], 2), np.int32)
    bt[:, 1] = times
    _, _, clusters, ncl = ops.cluster_peaks(to_dev(bt), 1, float(C0), float(C1))
    c = clusters[:int(ncl[0].item())].cpu().numpy()
    return tuple((times.dtype.type(a), int(b)) for a, b in c)
