"""Host-side multi-rank EM logic on CPU: two gloo ranks, each holding half the rows, must
produce the statistics / templates / log-likelihood trace of the single-rank run.  The local
E/M arithmetic is a NumPy stand-in with the engine interface (the CUDA engine needs a GPU);
what is under test is em_driver: sharding, the allreduce, the identical finalize."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpyEngine:
    def __init__(self, X, D, K):
        self.X = X.astype(np.float32)
        self.N, self.D, self.K = X.shape[0], D, K
        self.aff = torch.zeros((self.N, K), dtype=torch.float32)
        self.stats = torch.zeros(K * D + K, dtype=torch.float32)
        self.ll = torch.zeros(1, dtype=torch.float64)

    def new_params(self):
        z = lambda *s: torch.zeros(s, dtype=torch.float32)  # noqa: E731
        return dict(templates=z(self.K, self.D), logT=z(self.K, self.D), logInvT=z(self.K, self.D),
                    weights=z(self.K), logw=z(self.K))

    def set_affinities(self, aff):
        self.aff.copy_(torch.as_tensor(np.asarray(aff, np.float32)))

    def affinities(self):
        return self.aff

    def estep(self, params, hard, impl=0):
        S = self.X @ params["logT"].numpy().T + (1 - self.X) @ params["logInvT"].numpy().T
        ll = S.astype(np.float64) + params["logw"].numpy()[None]
        mx = ll.max(1, keepdims=True)
        p = np.exp(ll - mx)
        self.ll[0] = float((np.log(p.sum(1)) + mx[:, 0]).sum())
        self.aff.copy_(torch.as_tensor((p / p.sum(1, keepdims=True)).astype(np.float32)))
        self.stats.zero_()
        self.stats[self.K * self.D:] = self.aff.sum(0)

    def accumulate(self):
        self.stats[:self.K * self.D] += torch.as_tensor((self.aff.numpy().T @ self.X).reshape(-1))

    def finalize(self, stats, N_total, min_prob, params):
        K, D = self.K, self.D
        w = stats[K * D:].numpy() / N_total
        T = np.clip(stats[:K * D].numpy().reshape(K, D) / N_total / w[:, None], min_prob, 1 - min_prob)
        params["templates"].copy_(torch.as_tensor(T))
        params["logT"].copy_(torch.as_tensor(np.log(T)))
        params["logInvT"].copy_(torch.as_tensor(np.log(1 - T)))
        params["weights"].copy_(torch.as_tensor(w))
        params["logw"].copy_(torch.as_tensor(np.log(w)))


def _run(engine, N_total, K, lo, group, iters=5):
    from tsr_b200.em_driver import EmDriver
    drv = EmDriver(engine, N_total, group=group)
    aff = np.zeros((engine.N, K), np.float32)
    rows = np.arange(lo, lo + engine.N)
    rows = rows[rows < K * (N_total // K)]
    aff[rows - lo, rows % K] = 1
    engine.set_affinities(aff)
    drv.init_from_affinities(0.05)
    trace = [drv.expectation()]
    for _ in range(iters):
        drv.maximization(0.05)
        trace.append(drv.expectation())
    return trace, drv.params["templates"].numpy().copy(), drv.params["weights"].numpy().copy()


def _worker(rank, world, port, X, K, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tsr_b200.em_driver import shard_rows
    lo, hi = shard_rows(X.shape[0], rank, world)
    trace, T, w = _run(NumpyEngine(X[lo:hi], X.shape[1], K), X.shape[0], K, lo, None)
    if rank == 0:
        np.savez(out, trace=trace, T=T, w=w)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_rows_partition():
    from tsr_b200.em_driver import shard_rows
    for N, W in ((10, 3), (100000, 8), (5, 8), (0, 2)):
        r = [shard_rows(N, k, W) for k in range(W)]
        assert r[0][0] == 0 and r[-1][1] == N
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_em_matches_single_rank(tmp_path):
    from tsr_b200 import synth
    X, _, _, _ = synth.em_data(N=501, D=40, K=3, seed=9)
    K = 3
    single = _run(NumpyEngine(X, 40, K), 501, K, 0, False)
    out = str(tmp_path / "r0.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, X, K, out), nprocs=2, join=True)
    r = np.load(out)
    np.testing.assert_allclose(r["trace"], single[0], rtol=1e-6)
    np.testing.assert_allclose(r["T"], single[1], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r["w"], single[2], rtol=1e-5)
