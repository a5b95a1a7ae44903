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
    """Host stand-in with the engine interface of em_driver.GpuEmEngine: float64 statistics vector
    [K*D numerators | K responsibility sums | log-likelihood], stopping rule evaluated on the host here."""

    def __init__(self, X, D, K):
        self.X = X.astype(np.float64)
        self.N, self.D, self.K = X.shape[0], D, K
        self.aff = torch.zeros((self.N, K), dtype=torch.float32)
        self.stats = torch.zeros(K * D + K + 1, dtype=torch.float64)

    def new_params(self):
        z = lambda *s: torch.zeros(s, dtype=torch.float32)  # noqa: E731
        return dict(templates=z(self.K, self.D), logT=z(self.K, self.D), logInvT=z(self.K, self.D),
                    weights=z(self.K), logw=z(self.K))

    def set_affinities(self, aff):
        self.aff.copy_(torch.as_tensor(np.asarray(aff, np.float32)))

    def affinities(self):
        return self.aff

    def _q(self):
        return self.aff.numpy().astype(np.float64)

    def estep(self, params, hard, impl=0, ctl=None, prepared=False):
        if ctl is not None and ctl["done"]:
            return
        S = self.X @ params["logT"].numpy().astype(np.float64).T + (1 - self.X) @ params["logInvT"].numpy().astype(np.float64).T
        ll = S + params["logw"].numpy()[None].astype(np.float64)
        mx = ll.max(1, keepdims=True)
        p = np.exp(ll - mx)
        self.aff.copy_(torch.as_tensor((p / p.sum(1, keepdims=True)).astype(np.float32)))
        K, D = self.K, self.D
        self.stats[K * D:K * D + K] = torch.as_tensor(self._q().sum(0))
        self.stats[K * D + K] = float((np.log(p.sum(1)) + mx[:, 0]).sum())

    def accumulate(self, impl=0, from_affinities=False, ctl=None):
        if ctl is not None and ctl["done"]:
            return
        K, D = self.K, self.D
        q = self._q()
        self.stats[:K * D] = torch.as_tensor((q.T @ self.X).reshape(-1))
        if from_affinities:
            self.stats[K * D:K * D + K] = torch.as_tensor(q.sum(0))

    def loglik(self):
        return float(self.stats[self.K * self.D + self.K])

    def finalize(self, stats, N_total, min_prob, params, ctl=None):
        if ctl is not None and ctl["done"]:
            return
        K, D = self.K, self.D
        st = stats.numpy().astype(np.float64)
        w = (st[K * D:K * D + K] / N_total).astype(np.float32)
        T = np.clip((st[:K * D].reshape(K, D) / N_total).astype(np.float32) / w[:, None], min_prob, 1 - min_prob)
        params["templates"].copy_(torch.as_tensor(T))
        params["logT"].copy_(torch.as_tensor(np.log(T)))
        params["logInvT"].copy_(torch.as_tensor(np.log(1 - T)))
        params["weights"].copy_(torch.as_tensor(w))
        params["logw"].copy_(torch.as_tensor(np.log(w)))

    # stopping rule: the same arithmetic as em_control_kernel
    def control_init(self, tol, mode, min_iter, max_iter):
        return dict(done=False, it=0, trace=[], mode=mode, tol=tol, min_iter=min_iter, max_iter=max_iter, old=None)

    def control_step(self, c):
        if c["done"]:
            return
        ll = self.loglik()
        if not c["trace"]:
            c["old"] = ll / (1.0 - c["tol"]) - 10.0 if c["mode"] == 0 else -np.inf
        c["trace"].append(ll)
        if c["mode"] == 0:
            go = (abs((c["old"] - ll) / c["old"]) > c["tol"] and c["it"] < c["max_iter"]) or c["it"] < c["min_iter"]
        else:
            go = ll - c["old"] > c["tol"] and c["it"] < c["max_iter"]
        if not go:
            c["done"] = True
            return
        c["old"] = ll
        c["it"] += 1

    def control_read(self, c):
        return c["done"], c["it"], list(c["trace"])


def _init(engine, N_total, K, lo, group):
    from tsr_b200.em_driver import EmDriver
    drv = EmDriver(engine, N_total, group=group)
    aff = np.zeros((engine.N, K), np.float32)
    rows = np.arange(lo, lo + engine.N)
    rows = rows[rows < K * (N_total // K)]
    aff[rows - lo, rows % K] = 1
    engine.set_affinities(aff)
    drv.init_from_affinities(0.05)
    return drv


def _run(engine, N_total, K, lo, group, iters=5):
    drv = _init(engine, N_total, K, lo, group)
    trace = [drv.expectation()]
    for _ in range(iters):
        drv.maximization(0.05)
        trace.append(drv.expectation())
    # the speculative loop with the stopping rule (what run_EM uses): batches of 3 queued iterations
    drv2 = _init(engine, N_total, K, lo, group)
    it, tr2 = drv2.run(1e-4, mode="relative", min_iter=3, max_iter=50, min_prob=0.05, batch=3)
    return trace, drv.params["templates"].numpy().copy(), drv.params["weights"].numpy().copy(), it, tr2, drv2.params["templates"].numpy().copy()


def _worker(rank, world, port, X, K, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tsr_b200.em_driver import shard_rows
    lo, hi = shard_rows(X.shape[0], rank, world)
    trace, T, w, it, tr2, T2 = _run(NumpyEngine(X[lo:hi], X.shape[1], K), X.shape[0], K, lo, None)
    if rank == 0:
        np.savez(out, trace=trace, T=T, w=w, it=it, tr2=tr2, T2=T2)
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
    np.testing.assert_allclose(r["trace"], single[0], rtol=1e-9)
    np.testing.assert_allclose(r["T"], single[1], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r["w"], single[2], rtol=1e-5)
    # the queued loop stops at the same iteration with the same trace on 1 and 2 ranks, and matches a plain host loop
    assert int(r["it"]) == single[3] and 3 <= single[3] < 50
    np.testing.assert_allclose(r["tr2"], single[4], rtol=1e-9)
    np.testing.assert_allclose(r["T2"], single[5], rtol=1e-5, atol=1e-6)
    assert len(single[4]) == single[3] + 1
    np.testing.assert_array_equal(single[4][:6], single[0][:6])
    tr = np.array(single[4])
    assert abs((tr[-2] - tr[-1]) / tr[-2]) <= 1e-4 and abs((tr[-3] - tr[-2]) / tr[-3]) > 1e-4
