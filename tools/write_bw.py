#!/usr/bin/env python3
"""Write-only HBM bandwidth on this GPU (torch fill of a 4 GiB tensor, CUDA events), next to the copy bandwidth
MEASURED_PEAKS.json holds: the draw path writes 3 bytes per pixel and reads almost nothing."""
import torch
x = torch.empty(1 << 30, dtype=torch.int32, device="cuda")
y = torch.empty(1 << 30, dtype=torch.int32, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn, nbytes in (("fill_ (write only)", lambda: x.fill_(7), x.numel() * 4), ("zero_ (write only)", lambda: x.zero_(), x.numel() * 4),
                         ("copy_ (read + write)", lambda: y.copy_(x), 2 * x.numel() * 4), ("sum (read only)", lambda: x.sum(), x.numel() * 4)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%-22s %.3f ms  %.0f GB/s" % (name, best, nbytes / best / 1e6))
