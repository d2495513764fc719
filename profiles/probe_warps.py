import ctypes as C, sys, torch
sys.path.insert(0, ".")
from mchammer_b200 import native
lib = native.lib()
sink = torch.zeros(1, dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
for threads in (32, 64):
    for per_sm in (8, 10, 12, 14, 16, 20, 24, 32):
        if threads * per_sm > 1024: continue
        blocks, iters, rows, cols = 148 * per_sm, 1000, 50, 950
        best = 0.0
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            native.check(lib.mch_probe_sweep(blocks, threads, iters, 2, rows, cols, C.c_void_p(sink.data_ptr()), C.c_void_p(stream)), "probe")
            e1.record(); e1.synchronize()
            best = max(best, rows * cols * iters * blocks / (e0.elapsed_time(e1) * 1e-3))
        print(f"sweep-only no-store: threads {threads} x {per_sm} CTAs/SM ({threads*per_sm//32} warps/SM): {best:.3e} pair terms/s")
