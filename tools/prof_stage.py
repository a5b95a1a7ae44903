"""Run one pipeline stage a few times on BASELINE-size inputs (for ncu / timing).
   python tools/prof_stage.py code_parts|detect|spread_edges|spread_codes|all [--batch 1024] [--iters 3]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsr_b200 import ops, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("stage")
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--simt", action="store_true")
ap.add_argument("--impl-code", type=int, default=0, help="tsr_code_parts impl: 0 = 16-bit hi+lo planes, 2 = int8 form + float64 refinement")
ap.add_argument("--parts", type=int, default=100)
ap.add_argument("--planes", type=int, default=3, help="3 = INT8X2 (detect_tn_kernel, the pipeline default), 4 = INT8X3 (detect_tc_kernel)")
a = ap.parse_args()
impl = ops.IMPL_SIMT if a.simt else ops.IMPL_TENSOR
B = a.batch
lp, lip = synth.parts(a.parts)
parts = ops.PartsModel(lp, lip)
Tm, bgd = synth.templates(M=32, L=40, Fd=36, Kd=100)
W, c = synth.linear_filters(Tm, bgd)
bank = ops.FilterBank(W, cs=c, planes=a.planes)
bits = ops.pack_edges(torch.from_numpy(synth.fast_edge_batch(B)).cuda())
sb = ops.spread_edges(bits, 1, None, "line", 40)
impl_code = impl if a.simt else a.impl_code
codes, _, _ = ops.code_parts(sb, parts, 12, impl=impl_code)
fb = ops.spread_codes(codes, a.parts, 2, 2)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for it in range(a.iters):
    ev[0].record()
    if a.stage in ("spread_edges", "all"):
        ops.spread_edges(bits, 1, None, "line", 40)
    if a.stage in ("code_parts", "all"):
        ops.code_parts(sb, parts, 12, impl=impl_code)
    if a.stage in ("spread_codes", "all"):
        ops.spread_codes(codes, a.parts, 2, 2)
    if a.stage in ("detect", "all"):
        ops.detect(fb, bank, impl=impl)
    if a.stage == "detect_float":              # A8 / A9: real-valued detectors (bit-exact sequential fp32, SIMT)
        if it == 0:
            g = torch.Generator(device="cuda").manual_seed(9)
            Ff = torch.rand((64, 300, 320), device="cuda", generator=g)
            LFf = torch.randn((8, 40, 320), device="cuda", generator=g)
        ops.detect_float(Ff, LFf)
    if a.stage == "threshold_blocks":          # N4 + N3 front end: spectrogram -> packed edge bytes
        if it == 0:
            g = torch.Generator(device="cuda").manual_seed(1)
            spec = torch.randn((B, 40, 300), dtype=torch.float64, device="cuda", generator=g).cumsum(2) * 0.03
        ops.edges_from_spectrogram(spec, 40, 0.7)
    ev[1].record()
    torch.cuda.synchronize()
    print("%s iter %d: %.3f ms" % (a.stage, it, ev[0].elapsed_time(ev[1])))
