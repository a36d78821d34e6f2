#!/usr/bin/env python
"""Bootstrap kernel timings at full population size (development aid): count pass vs gather pass."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mcframework_b200 import _native as N, _stats_device as SD, _streams as S
from mcframework_b200._device import _ptr, _stream_ptr

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 200
lib = N.lib()
if len(sys.argv) > 3:
    print('L2 fetch granularity ->', sys.argv[3], 'rc', lib.mcf_set_l2_fetch_granularity(int(sys.argv[3])))
x = torch.from_numpy(np.random.default_rng(0).normal(5, 2, n)).cuda()
gen = np.random.default_rng(7)
kind, rec = S.generator_record(gen)
d_rec = torch.from_numpy(rec.view(np.uint8).copy()).cuda()
p = ((1 << 32) % n) / float(1 << 32)
total_slots = int(B * n / (1 - p) * 1.001) + 64
cs = 8192
nbytes = int(lib.mcf_bootstrap_workspace_bytes(total_slots, cs))
ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
total = torch.zeros(1, dtype=torch.int64, device="cuda")
sums = torch.zeros(B, dtype=torch.float64, device="cuda")
end_slot = torch.zeros(1, dtype=torch.int64, device="cuda")


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


for rep in range(2):
    sums.zero_()
    a = ev()
    N.check("count", lib.mcf_bootstrap_count(kind, _ptr(d_rec), n, 0, total_slots, cs, _ptr(ws), nbytes, _ptr(total), _stream_ptr()))
    b = ev()
    N.check("acc", lib.mcf_bootstrap_accumulate(kind, _ptr(d_rec), _ptr(x), n, B, 0, total_slots, cs, 0, _ptr(ws), nbytes, _ptr(sums),
                                                _ptr(end_slot), _stream_ptr()))
    c = ev()
    torch.cuda.synchronize()
    print(f"n={n} B={B}: count {a.elapsed_time(b):.1f} ms, accumulate {b.elapsed_time(c):.1f} ms "
          f"({B * n / b.elapsed_time(c) / 1e6:.2f} G gathers/s), accepted {int(total)} of {total_slots}")
