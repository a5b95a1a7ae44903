"""How does this box move pinned host memory to the GPU?  Call time vs device time of non_blocking copies."""
import time, torch
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
for mb in (12, 98):
    src = torch.empty(mb << 20, dtype=torch.uint8).pin_memory(); src.fill_(3)
    dst = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    s = torch.cuda.Stream()
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        with torch.cuda.stream(s):
            e0.record(s); dst.copy_(src, non_blocking=True); e1.record(s)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print("H2D %3d MB: call %.3f ms, until done %.3f ms, device time %.3f ms = %.1f GB/s" % (mb, (t1 - t0) * 1e3, (t2 - t0) * 1e3, e0.elapsed_time(e1), (mb << 20) / e0.elapsed_time(e1) / 1e6))
    out = torch.empty(2 << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(2 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize(); t0 = time.perf_counter(); out.copy_(d, non_blocking=True); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("D2H 2 MB: call %.3f ms, until done %.3f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
# concurrent compute + copy
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
src = torch.empty(12 << 20, dtype=torch.uint8).pin_memory(); dst = torch.empty(12 << 20, dtype=torch.uint8, device=dev)
s = torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    b = a @ a
    with torch.cuda.stream(s):
        dst.copy_(src, non_blocking=True)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("20 x (matmul + 12 MB H2D on a side stream): issue %.2f ms, total %.2f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
