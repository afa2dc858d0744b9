"""Pinned H2D / D2H bandwidth of the box (what bounds the host-pointer entry points)."""
import torch, time
n = 436 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, (a, b) in {"h2d": (d, h), "d2h": (h, d)}.items():
    for _ in range(2): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): a.copy_(b, non_blocking=True)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms:.3f} ms for {n/1e6:.0f} MB -> {n/ms/1e6:.1f} GB/s", flush=True)
