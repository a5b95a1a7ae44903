"""Time the E and M contractions of config 4 alone (debug experiments with TSR_DEBUG_BG_DBG)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tsr_b200 import ops
N, D, K = 100000, 12800, 16
g = torch.Generator(device="cuda").manual_seed(3)
rw = ops.rowwords(D)
X = torch.randint(-2**31, 2**31 - 1, (N, rw), dtype=torch.int32, device="cuda", generator=g)
X &= torch.randint(-2**31, 2**31 - 1, (N, rw), dtype=torch.int32, device="cuda", generator=g)
st = ops.EmState(X, D, K)
T = torch.rand((K, D), device="cuda", generator=g).clamp_(0.05, 0.95)
logT, logI, logw = torch.log(T), torch.log(1 - T), torch.full((K,), float(np.log(1.0 / K)), device="cuda")
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("dbg=%s  estep %.1f us  mstep %.1f us" % (os.environ.get("TSR_DEBUG_BG_DBG", "0"), t(lambda: st.estep(logT, logI, logw)), t(lambda: st.mstep_accumulate())))
