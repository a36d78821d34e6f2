#!/usr/bin/env python
"""How much faster is a random 8-byte gather when the table fits L2?  (design probe for a binned bootstrap)"""
import torch

def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

m = 200_000_000
for mb in (16, 32, 64, 96, 128, 256, 800):
    n = mb * 1024 * 1024 // 8
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    idx = torch.randint(0, n, (m,), device="cuda", dtype=torch.int32)
    out = torch.empty(m, dtype=torch.float64, device="cuda")
    ms = t(lambda: torch.index_select(x, 0, idx, out=out))
    print(f"table {mb:4d} MB: {m / ms / 1e6:7.1f} G gathers/s (torch.index_select, 4 B index read + 8 B write per gather)")
