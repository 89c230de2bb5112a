"""Diagnostic: PCIe D2H / H2D rates with pinned memory (torch copy engine), alone and with host threads streaming stores next to it."""
import time, torch, numpy as np
dev = torch.device("cuda", 0)
for mb in (1, 4, 20, 80, 320):
    n = mb << 20
    d = torch.zeros(n, dtype=torch.uint8, device=dev)
    h = torch.zeros(n, dtype=torch.uint8).pin_memory()
    for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{name} {mb:4d} MB pinned: {ms*1e3:8.1f} us  {n/ms/1e6:6.1f} GB/s", flush=True)
# a kernel writing straight into mapped pinned memory
n = 80 << 20
h = torch.zeros(n // 4, dtype=torch.int32).pin_memory()
d = torch.ones(n // 4, dtype=torch.int32, device=dev)
