"""Multi-GPU EM check (run under torchrun, one rank per GPU, NCCL): rows sharded with
em_driver.shard_rows, sufficient statistics all-reduced once per iteration.  Rank 0 compares the
log-likelihood trace / templates with a single-GPU run of the same problem and times an iteration
at BASELINE config 4 shapes.
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/em_dist_check.py"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from tsr_b200 import synth, ops
from tsr_b200.bernoulli_mixture import BernoulliMixture
from tsr_b200.em_driver import shard_rows, EmDriver

X, _, _, _ = synth.em_data(N=6000, D=640, K=8, seed=5)
lo, hi = shard_rows(X.shape[0], rank, world)
m = BernoulliMixture(8, X, row_range=(lo, hi))
m.run_EM(1e-5, min_probability=.01, max_iter=15)
if rank == 0:
    s = BernoulliMixture(8, X)                 # single-GPU reference on rank 0's device (no group)
    s.run_EM(1e-5, min_probability=.01, max_iter=15)
    n = min(len(s.loglikelihood_trace), len(m.loglikelihood_trace))
    rel = np.max(np.abs(np.array(s.loglikelihood_trace[:n]) - np.array(m.loglikelihood_trace[:n])) / np.abs(s.loglikelihood_trace[:n]))
    dT = np.abs(s.templates - m.templates).max()
    print("ranks %d: loglik trace rel diff %.2e, templates max abs diff %.2e, iterations %d vs %d" % (world, rel, dT, m.iterations, s.iterations))
    assert rel < 1e-5 and dT < 1e-3
dist.barrier()

# timing at config 4 shapes: N = 100 000 rows in total, sharded; synthetic bits generated on the device
N, D, K = 100000, 12800, 16
lo, hi = shard_rows(N, rank, world)
g = torch.Generator(device="cuda").manual_seed(3 + rank)
rw = ops.rowwords(D)
Xb = torch.randint(-2**31, 2**31 - 1, (hi - lo, rw), dtype=torch.int32, device="cuda", generator=g)
Xb &= torch.randint(-2**31, 2**31 - 1, (hi - lo, rw), dtype=torch.int32, device="cuda", generator=g)


class Eng:   # minimal engine around ops.EmState (the data is already packed on the device)
    def __init__(self):
        self.state = ops.EmState(Xb, D, K); self.N, self.D, self.K, self.device = hi - lo, D, K, Xb.device
    def new_params(self):
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=self.device)
        return dict(templates=z(K, D), logT=z(K, D), logInvT=z(K, D), weights=z(K), logw=z(K))
    def affinities(self): return self.state.aff
    def estep(self, p, hard, impl=0): self.state.estep(p["logT"], p["logInvT"], p["logw"], hard=hard, impl=impl)
    def accumulate(self): self.state.mstep_accumulate()
    stats = property(lambda self: self.state.stats)
    ll = property(lambda self: self.state.ll)
    def finalize(self, stats, Nt, mp, p): ops.em_finalize(stats, Nt, D, K, mp, p["templates"], p["logT"], p["logInvT"], p["weights"], p["logw"])


drv = EmDriver(Eng(), N, group=None)
T0 = torch.rand((K, D), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)).clamp_(0.05, 0.95)
drv.params["templates"].copy_(T0); drv.params["logT"].copy_(torch.log(T0)); drv.params["logInvT"].copy_(torch.log(1 - T0))
drv.params["weights"].fill_(1.0 / K); drv.params["logw"].fill_(float(np.log(1.0 / K)))
for _ in range(3):
    drv.expectation(); drv.maximization(0.05)
torch.cuda.synchronize(); dist.barrier()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 10
ev0.record()
for _ in range(iters):
    ll = drv.expectation(); drv.maximization(0.05)
ev1.record(); torch.cuda.synchronize()
t = torch.tensor([ev0.elapsed_time(ev1) / iters], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("config 4 (N=100000, D=12800, K=16) on %d GPUs: %.3f ms per EM iteration (E + M + allreduce + finalize), loglik %.6e" % (world, t.item(), ll))
dist.barrier()
dist.destroy_process_group()
