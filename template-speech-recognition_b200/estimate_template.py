"""Host-side weight preparation of the detection stage: drop-in for
``estimate_template.construct_linear_filter(s)`` (template_speech_rec/estimate_template.py:
156-172, 230-251).  Runs once per model on the host in float64 exactly like the reference and
produces the float32 filters / constants that ops.FilterBank uploads; it is model preparation,
not part of the per-utterance hot loop.
"""
import numpy as np


def construct_linear_filter(T, bgd, min_prob=.01, all_cs=False, use_spectral=False, sigma=None):
    """W = log((T/B)/((1-T)/(1-B))) as float32, c = sum log((1-T)/(1-B)) (or its running sum
    over time when all_cs).  T (L, ...) probabilities, bgd one frame (...), tiled over time."""
    T = np.clip(T, min_prob, 1 - min_prob)
    Bgd = np.tile(bgd, (T.shape[0],) + tuple(1 for _ in range(T.ndim - 1)))   # float reps in the reference
    T_inv = 1. - T
    Bgd_inv = 1. - Bgd
    C_exp_inv = T_inv / Bgd_inv
    if all_cs:
        c = np.cumsum(np.log(C_exp_inv.reshape(len(C_exp_inv), -1)).sum(1))
    else:
        c = np.log(C_exp_inv).sum()
    expW = (T / Bgd) / C_exp_inv
    return np.log(expW).astype(np.float32), c


def construct_linear_filters(Ts, bgd, all_cs=False, use_spectral=False, T_sigmas=None, bgd_sigma=None, min_prob=.01):
    """:156-172, Bernoulli branch (the spectral/Gaussian branch is outside the scoring path)."""
    if use_spectral:
        raise NotImplementedError("spectral (Gaussian) filters are outside the Bernoulli scoring path")
    return tuple(construct_linear_filter(np.clip(T, min_prob, 1 - min_prob), np.clip(bgd, min_prob, 1 - min_prob),
                                         all_cs=all_cs) for T in Ts)
