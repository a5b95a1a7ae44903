"""Device time of each pipeline kernel at the BASELINE shapes, host launch overhead excluded:
a few long fill kernels are queued first so the host runs ahead of the GPU, then `reps` calls of the
stage are timed between two CUDA events on the launching stream.
   python tools/time_kernels.py [stage ...] [--batch 1024] [--reps 10] [--planes 4]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsr_b200 import ops, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("stages", nargs="*", default=["spread_edges", "code_parts", "spread_codes", "detect", "reduce"])
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--planes", type=int, default=4)
ap.add_argument("--radius", type=int, default=1)
ap.add_argument("--impl-code", type=int, default=2, help="tsr_code_parts impl: 0 = 16-bit hi+lo planes, 2 = int8 form + float64 refinement")
ap.add_argument("--cold", action="store_true", help="flush the L2 (write + read pass) before every call and time each call with its own events")
a = ap.parse_args()
B = a.batch
lp, lip = synth.parts(100)
parts = ops.PartsModel(lp, lip)
Tm, bgd = synth.templates(M=32, L=40, Fd=36, Kd=100)
W, c = synth.linear_filters(Tm, bgd)
bank = ops.FilterBank(W, cs=c, planes=a.planes)
bits = ops.pack_edges(torch.from_numpy(synth.fast_edge_batch(B)).cuda())
sb = ops.spread_edges(bits, a.radius, None, "line", 40)
codes, _, _ = ops.code_parts(sb, parts, 12)
fb = ops.spread_codes(codes, 100, 2, 2)
scores = ops.detect(fb, bank)
spec = None
if "edges_fused" in a.stages or "edge_map" in a.stages or "threshold" in a.stages:
    g = torch.Generator(device="cuda").manual_seed(1)
    spec = (torch.randn((B, 40, 300), dtype=torch.float64, device="cuda", generator=g).cumsum(2) * 0.03)
    mags = ops.edge_map(spec)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
calls = {
    "spread_edges": (lambda: ops.spread_edges(bits, a.radius, None, "line", 40), 2.0 * 300 * 40 * B, "B"),
    "code_parts": (lambda: ops.code_parts(sb, parts, 12, impl=a.impl_code), 2.0 * 296 * 36 * 200 * 100 * B, "FLOP"),
    "spread_codes": (lambda: ops.spread_codes(codes, 100, 2, 2), 17.0 * 36 * 296 * B, "B"),
    "detect": (lambda: ops.detect(fb, bank, out=scores), 2.0 * 257 * 40 * 3600 * 32 * B, "FLOP"),
    "edges_fused": (lambda: ops.edges_from_spectrogram(spec, 40, 0.7), 9.0 * 300 * 40 * B, "B"),
    "edge_map": (lambda: ops.edge_map(spec), 72.0 * 300 * 40 * B, "B"),
    "threshold": (lambda: ops.threshold_edges(mags, 40, 0.7), 65.0 * 300 * 40 * B, "B"),
    "reduce": (lambda: ops.detect_reduce(scores, bank.cs, 40), (32 * 4 + 6.0) * 296 * B, "B"),
}
for name in a.stages:
    fn, work, unit = calls[name]
    for _ in range(3):
        fn()
    best = 1e9
    for trial in range(3):
        for _ in range(4):
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / a.reps)
    if a.cold:
        rd = torch.zeros(256 << 18, dtype=torch.int32, device="cuda")
        ts = []
        for _ in range(a.reps):
            flush.fill_(1)
            rd.sum()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        best = sorted(ts)[len(ts) // 2]
    rate = work / (best * 1e-3)
    print("%-13s %8.1f us   %s" % (name, best * 1e3, ("%.0f GB/s" % (rate / 1e9)) if unit == "B" else ("%.1f TFLOP/s" % (rate / 1e12))))
