"""Helpers shared by the NumPy-facing (reference-signature) modules: Cython-style argument
checks, host<->device staging and small LRU caches of prepared models."""
import hashlib
from collections import OrderedDict

import numpy as np
import torch

_CTYPE = {np.dtype(np.uint8): "unsigned char", np.dtype(np.float32): "float", np.dtype(np.float64): "double",
          np.dtype(np.int64): "long", np.dtype(np.int32): "int", np.dtype(np.uint16): "unsigned short"}


def typed(a, dtype, ndim, expected_name):
    """Mimic Cython typed-buffer argument checking (np.ndarray[ndim=.., dtype=..]): wrong
    dtype / ndim raise ValueError with Cython's wording; C-contiguity is normalised."""
    if not isinstance(a, np.ndarray):
        raise TypeError("Argument has incorrect type (expected numpy.ndarray, got %s)" % type(a).__name__)
    if a.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndim, a.ndim))
    if a.dtype != np.dtype(dtype):
        raise ValueError("Buffer dtype mismatch, expected '%s' but got '%s'"
                         % (expected_name, _CTYPE.get(a.dtype, str(a.dtype))))
    return np.ascontiguousarray(a)


def device():
    if not torch.cuda.is_available():
        from ._lib import TsrError
        raise TsrError("no CUDA device: template-speech-recognition_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device(), non_blocking=False)


def digest(*arrays):
    h = hashlib.blake2b(digest_size=16)
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str((a.shape, a.dtype.str)).encode())
        h.update(a.view(np.uint8).reshape(-1).data)
    return h.digest()


class Cache:
    def __init__(self, n=4):
        self.n, self.d = n, OrderedDict()

    def get(self, key, make):
        if key in self.d:
            self.d.move_to_end(key)
            return self.d[key]
        v = make()
        self.d[key] = v
        while len(self.d) > self.n:
            self.d.popitem(last=False)
        return v
