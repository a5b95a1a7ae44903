"""e2e diagnosis: ScoringPipeline.stream over packed pinned batches, 10 vs 30 vs 60 steps, with the per-phase host timers."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsr_b200 import ops, synth
from tsr_b200.pipeline import ScoringPipeline
torch.cuda.set_device(0)
B = 1024
lp, lip = synth.parts(100)
Tm, bgd = synth.templates(M=32, L=40, Fd=36, Kd=100)
W, c = synth.linear_filters(Tm, bgd)
pipe = ScoringPipeline(lp, lip, W, c)
X = synth.fast_edge_batch(B)
P = torch.from_numpy(np.packbits(X != 0, axis=-1, bitorder="little")[..., 0].copy()).pin_memory()
Xp = torch.from_numpy(X).pin_memory()
for feed, kw, name in ((P, {}, "packed"), (Xp, dict(pack="host", pack_threads=8), "host-pack"), (Xp, dict(pack="device"), "device-pack")):
    for steps in (5, 10, 30, 60):
        for _ in pipe.stream((feed for _ in range(5)), **kw):
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for best, arg in pipe.stream((feed for _ in range(steps)), **kw):
            float(best[0, 0])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = {k: round(v * 1e3 / steps, 3) for k, v in pipe.stream_stats.items() if k.endswith("_s")}
        print("%-11s %2d steps: %.3f ms/step = %.1f M frames/s  per-step host ms: %s" % (name, steps, dt * 1e3 / steps, B * 300 * steps / dt / 1e6, st))
