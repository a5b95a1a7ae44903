"""Drop-in for ``compute_likelihood_linear_filter``
(template_speech_rec/compute_likelihood_linear_filter.pyx): sliding linear-filter (Bernoulli
template log-likelihood-ratio) detection.  Same names, argument order, dtypes and output
shapes as the reference, including `detect`'s length-T output whose last L-1 entries are 0
(:157-158).  GPU only; see ``ops`` / ``pipeline`` for the batched device-resident API.
"""
import numpy as np

from . import ops
from ._np import Cache, device, digest, to_dev, typed

_bank_cache = Cache(4)


def _check_dims(*dims):
    if any(d > 65535 for d in dims):
        raise ValueError("dimension > 65535: the reference holds sizes in uint16 and silently wraps "
                         "(compute_likelihood_linear_filter.pyx:153-157); feed overlapping chunks")


def _bank(LF4):
    return _bank_cache.get(digest(LF4), lambda: ops.FilterBank(LF4, device=device()))


def detect(F, LF):
    """:135-176.  F uint8 (T,D1,D2) (nonzero = feature present), LF float32 (L,D1,D2).
    out[t] = sum over tau,f,k of [F[t+tau,f,k] != 0] * LF[tau,f,k] for t <= T-L, else 0."""
    F = typed(F, np.uint8, 3, "UINT_t")
    LF = typed(LF, np.float32, 3, "DTYPE_t")
    T, D1, D2 = F.shape
    _check_dims(T, D1, D2, LF.shape[0])
    if LF.shape[1:] != F.shape[1:]:
        raise ValueError("F %s and LF %s disagree on the feature dimensions" % (F.shape, LF.shape))
    if T == 0 or D1 * D2 == 0 or LF.shape[0] > T:
        return np.zeros(T, np.float32)
    if LF.shape[0] > 256:
        raise ValueError("template length > 256 frames is not supported")
    bank = _bank(LF[None])
    fb = ops.pack_features(to_dev(F.reshape(1, T, D1 * D2)))
    return ops.detect(fb, bank)[0, :, 0].cpu().numpy()


def _float_in(F, LF, nd):
    F = typed(F, np.float32, nd, "DTYPE_t")
    LF = typed(LF, np.float32, nd, "DTYPE_t")
    _check_dims(*F.shape, LF.shape[0])
    if LF.shape[1:] != F.shape[1:]:
        raise ValueError("F %s and LF %s disagree on the feature dimensions" % (F.shape, LF.shape))
    return F, LF


def _detect_float(F, LF):
    T, L = F.shape[0], LF.shape[0]
    D = int(np.prod(F.shape[1:]))
    if T == 0 or D == 0 or L > T:
        return np.zeros(T, np.float32)
    out = ops.detect_float(to_dev(F.reshape(1, T, D)), to_dev(LF.reshape(1, L, D)))
    return out[0, :, 0].cpu().numpy()


def detect_float(F, LF):
    """:178-219: real-valued features, out[t] = sum LF * F (float32, reference summation order)."""
    return _detect_float(*_float_in(F, LF, 3))


def detect_float2d(F, LF):
    """:281-318: 2-D (time, frequency) variant of detect_float."""
    return _detect_float(*_float_in(F, LF, 2))


def detect_spectral(F, LF):
    """:322-359: same arithmetic as detect_float2d (F * LF instead of LF * F)."""
    return _detect_float(*_float_in(F, LF, 2))


def detect_float_max(F, LF, cs):
    """:221-278: M filters with constants; returns (max_m score, uint16 argmax), first maximum wins;
    the L-1 tail entries of both outputs are 0."""
    F = typed(F, np.float32, 3, "DTYPE_t")
    LF = typed(LF, np.float32, 4, "DTYPE_t")
    cs = typed(cs, np.float32, 1, "DTYPE_t")
    T = F.shape[0]
    M, L = LF.shape[0], LF.shape[1]
    _check_dims(*F.shape, M, L)
    if LF.shape[2:] != F.shape[1:] or cs.shape[0] != M:
        raise ValueError("F %s, LF %s and cs %s disagree" % (F.shape, LF.shape, cs.shape))
    D = int(np.prod(F.shape[1:]))
    if T == 0 or D == 0 or L > T or M == 0:
        return np.zeros(T, np.float32), np.zeros(T, np.uint16)
    # each accumulation starts at cs[m] like the reference (:258), so the scores are bit-identical
    scores = ops.detect_float(to_dev(F.reshape(1, T, D)), to_dev(LF.reshape(M, L, D)), init=to_dev(cs))
    best, arg = ops.detect_reduce(scores, None, L, tail_mode=0)
    return best[0].cpu().numpy(), arg[0].cpu().numpy().view(np.uint16)


def _examples(F, LF, lengths):
    F = typed(F, np.uint8, 4, "UINT_t")
    LF = typed(LF, np.float64, 3, "float64_t")
    lengths = typed(lengths, np.int64, 1, "int64_t")
    n, T, D1, D2 = F.shape
    L = LF.shape[0]
    if LF.shape[1:] != F.shape[2:] or lengths.shape[0] != n:
        raise ValueError("F %s, LF %s and lengths %s disagree" % (F.shape, LF.shape, lengths.shape))
    if lengths.size and (lengths.min() < 0 or lengths.max() > T):
        raise ValueError("lengths must lie in [0, T]")
    if L > 256:
        raise ValueError("template length > 256 frames is not supported")
    # right-truncated filter (max_time = min(T - t, L), :116-117) == zero padding L-1 frames
    Fp = np.zeros((n, T + L - 1, D1 * D2), np.uint8)
    Fp[:, :T] = F.reshape(n, T, D1 * D2)
    bank = _bank(LF.astype(np.float32)[None])
    fb = ops.pack_features(to_dev(Fp))
    return ops.detect(fb, bank)[:, :T, 0].cpu().numpy().astype(np.float64), lengths


def detect_examples(F, LF, lengths, max_length):
    """:79-131 (intended semantics: the reference allocates a float32 array into a float64-typed
    buffer and raises).  F uint8 (n,T,D1,D2), LF float64 (L,D1,D2) -> float64 (n, max_length)."""
    s, lengths = _examples(F, LF, lengths)
    out = np.zeros((s.shape[0], int(max_length)), np.float64)
    for e, ln in enumerate(lengths):
        out[e, :ln] = s[e, :ln]
    return out


def get_max_detect_examples(F, LF, lengths):
    """:21-76: per example max over the first lengths[e] offsets (0.0 when lengths[e] == 0)."""
    s, lengths = _examples(F, LF, lengths)
    return np.array([s[e, :ln].max() if ln > 0 else 0.0 for e, ln in enumerate(lengths)], np.float64)
