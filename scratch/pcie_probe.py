"""Pinned host<->device copy bandwidth at the e2e transfer sizes (the ceiling of bench.py's e2e leg)."""
import torch
for mb in (2, 8, 16, 64):
    n = mb * (1 << 20) // 4
    h = torch.empty(n, dtype=torch.float32).pin_memory(); d = torch.empty(n, dtype=torch.float32, device="cuda")
    h2 = torch.empty(n // 4, dtype=torch.float32).pin_memory(); d2 = torch.empty(n // 4, dtype=torch.float32, device="cuda")
    s2 = torch.cuda.Stream()
    for both in (False, True):
        for _ in range(3): d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            d.copy_(h, non_blocking=True)
            if both:
                with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        print("H2D %3d MB%s: %.1f GB/s" % (mb, " (+ concurrent D2H of a quarter)" if both else "", 20 * n * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9))
