"""Drop-in for ``cluster_times.cluster_times`` (template_speech_rec/cluster_times.pyx:133-219), the clustering of
detection peaks that follows the peak picking.  NumPy in, NumPy out; runs ``tsr_cluster_peaks`` on the GPU (one thread
per utterance -- the batched form is ``ops.cluster_peaks``); there is no CPU path."""
import numpy as np

from . import ops
from ._np import to_dev, typed


def cluster_times(times, C0, C1):
    """times uint16 (n,), strictly increasing peak frames -> cluster_partition uint8 (n+1,): 2 marks the first peak of
    a cluster (and the closing entry n), 0 a continuation."""
    times = typed(times, np.uint16, 1, "npy_uint16")
    n = times.shape[0]
    bt = np.zeros((n, 2), np.int32)
    bt[:, 1] = times
    _, part, _, _ = ops.cluster_peaks(to_dev(bt), 1, float(C0), float(C1))
    return part.cpu().numpy()
