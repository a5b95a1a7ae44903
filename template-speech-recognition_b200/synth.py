"""Seeded synthetic inputs for tests and bench.  NumPy only.

Utterance ``u`` uses seed ``1000+u``: the time axis is cut into segments of length
U{20..80} that alternate "silence" (p=0.02) and "speech" (p=0.20); every edge bit is an
independent Bernoulli(p).  Overall density is about 0.11, so 5x5x8 windows fall on both
sides of the edge-count threshold (12 by default, main_multiprocessing.py:71).
"""
import numpy as np


def utterance_edges(u, T=300, F=40, E=8):
    rs = np.random.RandomState(1000 + u)
    p = np.empty(T)
    t, speech = 0, bool(rs.randint(2))
    while t < T:
        n = rs.randint(20, 81)
        p[t:t + n] = 0.20 if speech else 0.02
        speech = not speech
        t += n
    return (rs.rand(T, F, E) < p[:, None, None]).astype(np.uint8)


def utterance_batch(B, T=300, F=40, E=8, first=0):
    return np.stack([utterance_edges(first + u, T, F, E) for u in range(B)])


def fast_edge_batch(B, T=300, F=40, E=8, seed=7):
    """Same two-density segment structure, generated in one vectorised draw (for the
    large bench batches where per-utterance RandomState construction dominates)."""
    rs = np.random.RandomState(seed)
    nseg = T // 20 + 2
    lens = rs.randint(20, 81, size=(B, nseg))
    start = rs.randint(2, size=(B, 1))
    seg_speech = (np.arange(nseg)[None, :] + start) % 2
    ends = np.cumsum(lens, axis=1)
    t = np.arange(T)
    seg_id = (t[None, :, None] >= ends[:, None, :]).sum(-1)
    p = np.where(np.take_along_axis(seg_speech, seg_id, axis=1) == 1, 0.20, 0.02)
    return (rs.rand(B, T, F, E) < p[:, :, None, None]).astype(np.uint8)


def parts(K=100, ph=5, pw=5, E=8, seed=1):
    """log_parts, log_invparts as float64 (what code_parts requires: code_parts.pyx:71-72)."""
    rs = np.random.RandomState(seed)
    P = np.clip(rs.rand(K, ph, pw, E) ** 2, .01, .99)
    return np.log(P), np.log(1 - P)


def templates(M=32, L=40, Fd=36, Kd=100, seed=2):
    """Mixture templates (M,L,Fd,Kd) and one-frame background (Fd,Kd)."""
    rs = np.random.RandomState(seed)
    T = np.clip(rs.rand(M, L, Fd, Kd) ** 2, .01, .99)
    bgd = np.clip(0.3 * rs.rand(Fd, Kd), .01, .4)
    return T, bgd


def linear_filters(T, bgd):
    """(W float32 (M,L,Fd,Kd), c float32 (M,)) as estimate_template.construct_linear_filter
    (estimate_template.py:230-251) would build them."""
    T = np.clip(T, .01, .99)
    Tinv, Binv = 1. - T, 1. - bgd
    W = np.log((T / bgd) / (Tinv / Binv)).astype(np.float32)
    c = np.log(Tinv / Binv).reshape(T.shape[0], -1).sum(1)
    return W, c.astype(np.float32)


def em_data(N=100000, D=12800, K=16, seed=3, dtype=np.uint8, chunk=8192):
    """Planted Bernoulli mixture: K templates clip(U^2,.05,.95), weights ~ Dirichlet(1)."""
    rs = np.random.RandomState(seed)
    T = np.clip(rs.rand(K, D) ** 2, .05, .95).astype(np.float32)
    w = rs.dirichlet(np.ones(K))
    z = rs.choice(K, size=N, p=w)
    X = np.empty((N, D), dtype)
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        X[s:e] = rs.rand(e - s, D).astype(np.float32) < T[z[s:e]]
    return X, z, T, w
