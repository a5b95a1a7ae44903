"""Time one EM iteration (E-step + M-step accumulation + finalize) at BASELINE config 4 shapes.
   python tools/bench_em.py [--N 100000] [--D 12800] [--K 16] [--iters 5]"""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsr_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=100000)
ap.add_argument("--D", type=int, default=12800)
ap.add_argument("--K", type=int, default=16)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
N, D, K = a.N, a.D, a.K
rw = ops.rowwords(D)
g = torch.Generator(device="cuda").manual_seed(3)
X = torch.randint(-2**31, 2**31 - 1, (N, rw), dtype=torch.int32, device="cuda", generator=g)
X &= torch.randint(-2**31, 2**31 - 1, (N, rw), dtype=torch.int32, device="cuda", generator=g)   # ~25 % ones
st = ops.EmState(X, D, K)
T = torch.rand((K, D), device="cuda").clamp_(0.05, 0.95)
logT, logI = torch.log(T), torch.log(1 - T)
logw = torch.full((K,), float(np.log(1.0 / K)), device="cuda")
params = dict(templates=T.clone(), logT=logT, logInvT=logI, weights=torch.full((K,), 1.0 / K, device="cuda"), logw=logw)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for it in range(a.iters):
    ev[0].record()
    st.estep(params["logT"], params["logInvT"], params["logw"])
    ev[1].record()
    st.mstep_accumulate()
    ev[2].record()
    ops.em_finalize(st.stats, N, D, K, 0.05, params["templates"], params["logT"], params["logInvT"], params["weights"], params["logw"])
    ev[3].record()
    torch.cuda.synchronize()
    print("iter %d: E %.3f ms  M %.3f ms  finalize %.3f ms  total %.3f ms  ll %.6e" % (
        it, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), ev[0].elapsed_time(ev[3]), float(st.ll.item())))
print("X bytes %.1f MB -> HBM floor %.1f us per pass" % (X.numel() * 4 / 1e6, X.numel() * 4 / 6541.8e9 * 1e6))
