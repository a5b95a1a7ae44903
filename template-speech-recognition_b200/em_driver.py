"""Host-side driver of Bernoulli-mixture EM over row shards (one shard per rank / GPU).

The E- and M-step arithmetic lives behind a small "local engine" interface so that the
host logic -- sharding, the allreduce of the sufficient statistics, the speculative
iteration loop -- is the same code whether the engine is the CUDA one (``GpuEmEngine``,
the product) or a stand-in used by the CPU/gloo tests.

Per EM iteration and rank:   stopping rule on the reduced log-likelihood (``control_step``)
->  finalize (divide, clip, log: identical on every rank)  ->  E-step on the local rows
(posteriors, responsibility sums, log-likelihood)  ->  M accumulation (numerators)  ->  ONE
allreduce(sum) of the statistics vector [K*D numerators | K responsibility sums |
log-likelihood] (float64).  Inside a rank the CUDA engine forms every entry as an exact integer
sum, so a rank's statistics do not depend on the order of summation.  The stopping rule runs on
the device: the host queues ``batch`` iterations at a time and reads the control block back once
per batch; iterations past the stop are no-ops.  With a single
rank no collective is issued.  Scoring/coding/detection never communicate.
"""
import os

import numpy as np
import torch


def shard_rows(N, rank, world):
    """Contiguous, balanced row range of `rank` out of `world` (first N % world ranks get one more)."""
    base, extra = divmod(N, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class GpuEmEngine:
    """Local engine backed by libtsr_b200.so: bit-packed rows resident in HBM."""

    def __init__(self, data_rows, D, K, device=None, chunk_rows=16384, packed=None):
        from . import ops
        from ._np import device as _device
        self.ops = ops
        self.device = _device() if device is None else device
        if packed is not None:                  # rows already bit-packed on the device
            X = packed
            n = X.shape[0]
        else:
            n = data_rows.shape[0]
            rw = ops.rowwords(D)
            X = torch.empty((n, rw), dtype=torch.int32, device=self.device)
            for s in range(0, n, chunk_rows):      # stage through uint8 chunks; pack on the device
                blk = np.ascontiguousarray(data_rows[s:s + chunk_rows])
                blk = (blk != 0).astype(np.uint8) if blk.dtype != np.uint8 else blk
                d = torch.from_numpy(blk.reshape(1, -1, D)).to(self.device)
                X[s:s + blk.shape[0]] = ops.pack_features(d)[0]
        self.state = ops.EmState(X, D, K)
        self.N, self.D, self.K = n, D, K
        self.graph_ok = True                    # every launch of an iteration is stream ordered and capturable

    def new_params(self):
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=self.device)  # noqa: E731
        return dict(templates=z(self.K, self.D), logT=z(self.K, self.D), logInvT=z(self.K, self.D),
                    weights=z(self.K), logw=z(self.K))

    def set_affinities(self, aff):
        self.state.aff.copy_(torch.as_tensor(np.asarray(aff, dtype=np.float32)))

    def affinities(self):
        return self.state.aff

    @property
    def stats(self):
        return self.state.stats

    def estep(self, params, hard, impl=0, ctl=None, prepared=False, stats=None):
        """Posteriors for the local rows; fills the responsibility sums and the log-likelihood of stats.
        prepared: the weight planes were laid out by the last finalize (skip the per-call preparation)."""
        lt, li = (None, None) if (prepared and impl == 0) else (params["logT"], params["logInvT"])
        self.state.estep(lt, li, params["logw"], hard=hard, impl=impl, ctl=ctl, stats=stats)

    def accumulate(self, impl=0, from_affinities=False, ctl=None, stats=None):
        """stats[:K*D] = aff^T X over the local rows (from_affinities: from self.affinities(), also setting the
        responsibility sums -- used to initialise templates from a hard assignment)."""
        self.state.mstep_accumulate(impl=impl, aff=self.state.aff if from_affinities else None, ctl=ctl, stats=stats)

    def peer_stats(self, group):
        """Peer-mapped home of the statistics vector (NVLink one-shot allreduce), or raises."""
        return self.ops.PeerStats(self.state.stats.numel(), self.device, group)

    def loglik(self):
        return self.state.loglik()

    def finalize(self, stats, N_total, min_prob, params, ctl=None):
        self.ops.em_finalize(stats, N_total, self.D, self.K, min_prob, params["templates"], params["logT"],
                             params["logInvT"], params["weights"], params["logw"], state=self.state, ctl=ctl)

    # ---- stopping rule on the device
    def control_init(self, tol, mode, min_iter, max_iter):
        return self.ops.EmControl(self.device, tol, mode, min_iter, max_iter, max(min_iter, max_iter) + 2)

    def control_step(self, ctl):
        ctl.step(self.state.stats, self.D, self.K)

    def control_read(self, ctl):
        return ctl.read()


class EmDriver:
    def __init__(self, engine, N_total, group=None):
        self.e = engine
        self.N_total = int(N_total)
        self.group = group
        self.params = engine.new_params()
        self.dist = group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()
                                          and torch.distributed.get_world_size() > 1 and group is not False)
        if group is False:
            self.dist = False
        self.allreduces = 0
        self.graph_iters = 1
        # Sharded rows on GPUs of one NVLink domain: the sum of the statistics is ONE kernel of this library over peer
        # memory (ops.PeerStats / tsr_em_allreduce_peer) instead of an NCCL call; NCCL stays the path for engines or
        # systems without symmetric memory (and the CPU/gloo tests).  TSR_EM_PEER=0 forces NCCL.
        self.peer, self.peer_error, self.it = None, None, 0
        if self.dist and hasattr(engine, "peer_stats") and os.environ.get("TSR_EM_PEER", "1") != "0":
            ok = 0
            try:
                self.peer = engine.peer_stats(self.group if self.group not in (None, False) else None)
                ok = 1
            except Exception as exc:  # noqa: BLE001 -- no symmetric memory here: NCCL
                self.peer_error = "%s: %s" % (type(exc).__name__, exc)
            flag = torch.tensor([ok], dtype=torch.int32, device=engine.device)
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN,
                                         group=self.group if self.group not in (None, False) else None)
            if int(flag.item()) == 0:          # all ranks or none
                self.peer = None
            elif self.peer.two_phase:          # peers store the sums straight into this rank's statistics vector
                self.peer.out.copy_(engine.state.stats)
                engine.state.stats = self.peer.out

    def _src(self):
        """Where this iteration's local statistics go: the engine's vector, or the peer-mapped slot of this parity."""
        return None if self.peer is None else self.peer.src[self.it & 1]

    def _allreduce(self, ctl=None):
        """Global statistics -> engine.stats on every rank (the ONE collective of an iteration)."""
        if self.peer is not None:
            self.peer.allreduce(self.it & 1, self.e.stats, ctl)
            self.it += 1
            self.allreduces += 1
        elif self.dist:
            torch.distributed.all_reduce(self.e.stats, op=torch.distributed.ReduceOp.SUM,
                                         group=self.group if self.group not in (None, False) else None)
            self.allreduces += 1

    def _estep(self, hard, impl, ctl, prepared, accumulate=True):
        kw = {} if self.peer is None else {"stats": self._src()}
        self.e.estep(self.params, hard, impl, ctl=ctl, prepared=prepared, **kw)
        if accumulate:
            self.e.accumulate(impl, ctl=ctl, **kw)
        elif self.peer is not None:
            self.peer.src[self.it & 1][:self.e.K * self.e.D].zero_()      # no numerators this time: the slot's are stale
        self._allreduce(ctl)

    def expectation(self, hard=False, impl=0, accumulate=True):
        """E-step (+ numerator accumulation) on every rank, then the sufficient-statistics
        allreduce.  Returns the global log-likelihood (python float; synchronises)."""
        self._estep(hard, impl, None, False, accumulate)
        return self.e.loglik()

    def maximization(self, min_prob):
        """Finalise templates / weights from the reduced statistics (identical on every rank)."""
        self.e.finalize(self.e.stats, self.N_total, float(min_prob), self.params)

    def _iteration(self, hard, impl, ctl, min_prob):
        """One queued EM iteration: stopping rule, M-step proper, E-step, M accumulation, allreduce (no host sync)."""
        e = self.e
        e.control_step(ctl)                                     # continue? (:144)  -- every rank sees the same reduced value
        e.finalize(e.stats, self.N_total, float(min_prob), self.params, ctl=ctl)    # M-step (:150)
        self._estep(hard, impl, ctl, True)                      # E-step (:152)

    def capture_iteration(self, hard, impl, ctl, min_prob):
        """The ~11 launches (and the allreduce) of an iteration as ONE CUDA graph, or None when the engine cannot be
        captured: at 0.3 ms of GPU work per iteration the per-launch host cost would otherwise set the pace."""
        import os
        if not getattr(self.e, "graph_ok", False) or os.environ.get("TSR_EM_NO_GRAPH"):
            return None
        # the peer allreduce alternates between two source slots: the graph then holds an even number of iterations
        self.graph_iters = 2 if self.peer is not None else 1
        try:
            if self.peer is not None and (self.it & 1):
                self._iteration(hard, impl, ctl, min_prob)      # start the captured pair on an even launch
            side = torch.cuda.Stream(device=self.e.device)
            side.wait_stream(torch.cuda.current_stream(self.e.device))
            with torch.cuda.stream(side):
                for _ in range(self.graph_iters):
                    self._iteration(hard, impl, ctl, min_prob)  # warm-up on the capture stream (real iterations)
            torch.cuda.current_stream(self.e.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(self.graph_iters):
                    self._iteration(hard, impl, ctl, min_prob)
            return g
        except Exception:  # noqa: BLE001 -- e.g. a collective that cannot be captured: fall back to eager launches
            torch.cuda.synchronize()
            return None

    def run(self, tol, mode="relative", min_iter=10, max_iter=500, hard=False, min_prob=0.05, batch=8, impl=0):
        """The EM loop of run_EM (bernoulli_mixture.py:135-156; mode "absolute": bernoulli_em.py:22-34) with the
        stopping rule evaluated on the device.  Returns (iterations, log-likelihood trace)."""
        e = self.e
        ctl = e.control_init(float(tol), 0 if mode == "relative" else 1, int(min_iter), int(max_iter))
        self._estep(hard, impl, ctl, False)                         # first E-step (:135)
        graph = self.capture_iteration(hard, impl, ctl, min_prob)   # (runs one real iteration while warming up)
        while True:
            for _ in range(max(1, int(batch)) // (self.graph_iters if graph is not None else 1) or 1):
                if graph is not None:
                    graph.replay()
                else:
                    self._iteration(hard, impl, ctl, min_prob)
            done, iterations, trace = e.control_read(ctl)           # one host synchronisation per batch
            if done:
                return iterations, trace

    def init_from_affinities(self, min_prob, uniform_weights=True):
        """Templates = affinity-weighted data means (the deterministic round-robin 'unif_rand'
        start of bernoulli_mixture.py:195-237); weights stay uniform as in the reference."""
        K = self.e.K
        kw = {} if self.peer is None else {"stats": self._src()}
        self.e.accumulate(from_affinities=True, **kw)
        self._allreduce()
        self.e.finalize(self.e.stats, self.N_total, float(min_prob), self.params)
        if uniform_weights:
            self.params["weights"].fill_(1.0 / K)
            self.params["logw"].fill_(float(np.log(np.float32(1.0 / K))))
