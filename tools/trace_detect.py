"""Dump the per-stage timeline of CTA 0 of detect_tc_kernel (debug hook).
Needs a library built with the stamps compiled in:
    TSR_NVCC_EXTRA=-DTSR_DEBUG_TRACE python template-speech-recognition_b200/build.py --force
(the default build leaves them out: even predicated off they slow the issue loops by 2-7 %)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsr_b200 import ops, synth, _lib
lib = _lib.lib()
planes = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rs = np.random.RandomState(0)
if os.environ.get("TRACE_EM"):      # the EM E-step shape: N = 100 000 rows, D = 12 800 bits, K = 16 "templates" of one frame
    bank = ops.FilterBank((rs.randn(16, 1, 12800) * 0.5).astype(np.float32), planes=planes)
    g = torch.Generator(device="cuda").manual_seed(0)
    fb = torch.randint(-2**31, 2**31 - 1, (1, 100000, 400), dtype=torch.int32, device="cuda", generator=g)
else:
    Tm, bgd = synth.templates(M=32, L=40, Fd=36, Kd=100)
    W, c = synth.linear_filters(Tm, bgd)
    bank = ops.FilterBank(W, cs=c, planes=planes)
    fb = torch.from_numpy(rs.randint(-2**31, 2**31 - 1, (1024, 296, 114), dtype=np.int64).astype(np.int32)).cuda()
    fb &= torch.from_numpy(rs.randint(-2**31, 2**31 - 1, (1024, 296, 114), dtype=np.int64).astype(np.int32)).cuda()
ops.detect(fb, bank)
torch.cuda.synchronize()
buf = torch.zeros(256 * 8, dtype=torch.int64, device="cuda")
lib.tsr_debug_set_detect_trace.argtypes = [ctypes.c_void_p]
if len(sys.argv) > 2:
    lib.tsr_debug_set_detect_flags(int(sys.argv[2]))
    print("debug flags", sys.argv[2])
lib.tsr_debug_set_detect_trace(buf.data_ptr())
ops.detect(fb, bank)
torch.cuda.synchronize()
lib.tsr_debug_set_detect_trace(None)
t = buf.cpu().numpy().reshape(256, 8)
t0 = t[t > 0].min()
names = ["exp:empty_ok", "exp:arrived", "ld:empty_ok", "peer:relayed", "mma:full_ok", "mma:commit", "epi:tfull", "epi:done"]
print("planes", planes)
print("iter " + " ".join("%13s" % n for n in names))
for i in list(range(0, 4)) + list(range(60, 76)):
    print("%4d " % i + " ".join("%13d" % (x - t0 if x else -1) for x in t[i]))
d = np.diff(t[60:250, 5])
print("mean cycles between consecutive stage commits:", d.mean())
c5 = t[:, 5]
d = np.diff(c5[:250])
big = np.argsort(-d)[:12]
print("largest gaps between consecutive commits (iter: cycles):", sorted([(int(i), int(d[i])) for i in big]))
print("median stage period", np.median(d), "sum", d.sum())
e6, e7 = t[:8, 6], t[:8, 7]
print("epilogue (tfull -> done) per step:", [(int(b - a)) for a, b in zip(e6, e7) if a and b])
print("step starts:", [int(x - t0) for x in e6 if x])

tp = t[192:192 + 14, :6]
if tp.any():
    base = e6[1]
    print("drain passes of step 1 (cycles after tfull): ld_issued ld_done stored bar1 summed bar2")
    for row in tp:
        if row[0]:
            print("   " + " ".join("%7d" % (x - base) for x in row))
