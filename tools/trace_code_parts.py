"""Dump the per-tile timeline of CTA 0 of code_parts_tc_kernel (debug hook).
Needs a library built with the stamps compiled in:
    TSR_NVCC_EXTRA=-DTSR_DEBUG_TRACE python template-speech-recognition_b200/build.py --force
(the default build leaves them out: even predicated off they slow the issue loops by 2-7 %).
    python tools/trace_code_parts.py [flags] [impl]     impl: 0 = 16-bit hi+lo planes, 2 = int8 form (default)
flags (tsr_debug_set_code_flags): 1 = the epilogue skips the TMEM reads and the arg-max, 2 = the issuers skip the MMAs,
4 = epilogue detail stamps, 8 = int8 form without its float64 refinement launch.  Point TSR_B200_LIB at the trace build
to keep the product library untouched."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsr_b200 import ops, synth, _lib
lib = _lib.lib()
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lp, lip = synth.parts(100)
parts = ops.PartsModel(lp, lip)
bits = ops.pack_edges(torch.from_numpy(synth.fast_edge_batch(1024)).cuda())
ops.code_parts(bits, parts, 12, impl=impl)
torch.cuda.synchronize()
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
lib.tsr_debug_set_code_flags(flags)
print("debug flags", flags)
buf = torch.zeros(96 * 8, dtype=torch.int64, device="cuda")
lib.tsr_debug_set_code_trace.argtypes = [ctypes.c_void_p]
lib.tsr_debug_set_code_trace(buf.data_ptr())
ops.code_parts(bits, parts, 12, impl=impl)
torch.cuda.synchronize()
lib.tsr_debug_set_code_trace(None)
t = buf.cpu().numpy().reshape(96, 8)
t0 = t[t > 0].min()
names = ["prod:empty_ok", "prod:arrived", "mma:full_ok", "mma:tempty_ok", "mma:issued", "epi:tfull_ok", "epi:tmem_free", "epi:done"]
if flags & 4:   # epilogue detail: slots 1-4 = first wait::ld done, first block reduced, second wait done, second block reduced
    names = ["-", "epi:ld1_done", "epi:red1_done", "epi:ld2_done", "epi:red2_done", "epi:tfull_ok", "epi:tmem_free", "epi:done"]
print("tile " + " ".join("%14s" % n for n in names))
for i in range(30, 38):
    print("%4d " % i + " ".join("%14d" % (x - t0 if x else -1) for x in t[i]))
d = np.diff(t[8:90, 4])
print("mean cycles between consecutive tiles' MMA issue:", d.mean())
