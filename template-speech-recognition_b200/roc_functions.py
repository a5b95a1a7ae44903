"""Drop-in for the peak picking of the reference's ``roc_functions`` module
(template_speech_rec/roc_functions.py): the step right after detection.  NumPy in, NumPy out;
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
