"""Host-side driver of Bernoulli-mixture EM over row shards (one shard per rank / GPU).

The E- and M-step arithmetic lives behind a small "local engine" interface so that the
host logic -- sharding, the allreduce of the sufficient statistics, the stopping rules of
the two reference classes -- is the same code whether the engine is the CUDA one
(``GpuEmEngine``, the product) or a stand-in used by the CPU/gloo tests.

Per EM iteration and rank:   E-step on the local rows (posteriors, responsibility sums,
log-likelihood)  ->  M accumulation (numerators)  ->  ONE allreduce(sum) of the float32
vector [K*D numerators | K responsibility sums] plus one of the float64 log-likelihood
->  every rank finalises identical templates (divide, clip, log).  With a single rank no
collective is issued.  Scoring/coding/detection never communicate.
"""
import numpy as np
import torch


def shard_rows(N, rank, world):
    """Contiguous, balanced row range of `rank` out of `world` (first N % world ranks get one more)."""
    base, extra = divmod(N, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class GpuEmEngine:
    """Local engine backed by libtsr_b200.so: bit-packed rows resident in HBM."""

    def __init__(self, data_rows, D, K, device=None, chunk_rows=16384):
        from . import ops
        from ._np import device as _device
        self.ops = ops
        self.device = _device() if device is None else device
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

    def new_params(self):
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=self.device)  # noqa: E731
        return dict(templates=z(self.K, self.D), logT=z(self.K, self.D), logInvT=z(self.K, self.D),
                    weights=z(self.K), logw=z(self.K))

    def set_affinities(self, aff):
        self.state.aff.copy_(torch.as_tensor(np.asarray(aff, dtype=np.float32)))

    def affinities(self):
        return self.state.aff

    def estep(self, params, hard, impl=0):
        """Posteriors for the local rows; fills stats[K*D:] (responsibility sums) and ll."""
        self.state.estep(params["logT"], params["logInvT"], params["logw"], hard=hard, impl=impl)

    def accumulate(self):
        """stats[:K*D] += aff^T X over the local rows."""
        self.state.mstep_accumulate()

    @property
    def stats(self):
        return self.state.stats

    @property
    def ll(self):
        return self.state.ll

    def finalize(self, stats, N_total, min_prob, params):
        self.ops.em_finalize(stats, N_total, self.D, self.K, min_prob, params["templates"], params["logT"],
                             params["logInvT"], params["weights"], params["logw"])


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

    def _allreduce(self, t):
        if self.dist:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM,
                                         group=self.group if self.group not in (None, False) else None)

    def expectation(self, hard=False, impl=0, accumulate=True):
        """E-step (+ numerator accumulation) on every rank, then the sufficient-statistics
        allreduce.  Returns the global log-likelihood (python float)."""
        self.e.estep(self.params, hard, impl)
        if accumulate:
            self.e.accumulate()
        self._allreduce(self.e.stats)
        self._allreduce(self.e.ll)
        return float(self.e.ll.item())

    def maximization(self, min_prob):
        """Finalise templates / weights from the reduced statistics (identical on every rank)."""
        self.e.finalize(self.e.stats, self.N_total, float(min_prob), self.params)

    def init_from_affinities(self, min_prob, uniform_weights=True):
        """Templates = affinity-weighted data means (the deterministic round-robin 'unif_rand'
        start of bernoulli_mixture.py:195-237); weights stay uniform as in the reference."""
        K = self.e.K
        self.e.stats.zero_()
        self.e.accumulate()
        self.e.stats[K * self.e.D:] = self.e.affinities().sum(0)
        self._allreduce(self.e.stats)
        self.e.finalize(self.e.stats, self.N_total, float(min_prob), self.params)
        if uniform_weights:
            self.params["weights"].fill_(1.0 / K)
            self.params["logw"].fill_(float(np.log(np.float32(1.0 / K))))
