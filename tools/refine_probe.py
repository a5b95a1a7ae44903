"""Debug: how many locations the int8 coding kernel marks for the float64 refinement, and what the refinement costs."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsr_b200 import ops, synth, _lib
lib = _lib.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lp, lip = synth.parts(100)
parts = ops.PartsModel(lp, lip)
bits = ops.pack_edges(torch.from_numpy(synth.fast_edge_batch(B)).cuda())
sb = ops.spread_edges(bits, 1, None, "line", 40)
def timed(impl, K=100, reps=10):
    for _ in range(3):
        ops.code_parts(sb, parts, 12, impl=impl)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.code_parts(sb, parts, 12, impl=impl)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for K in (100, 200):
    lp, lip = synth.parts(K)
    parts = ops.PartsModel(lp, lip)
    lib.tsr_debug_set_code_flags(8)
    t_main = timed(2)
    lib.tsr_debug_set_code_flags(0)
    print("K=%d: hi+lo planes %.1f us | int8 main kernel %.1f us, with refinement %.1f us" % (K, timed(0), t_main, timed(2)))
