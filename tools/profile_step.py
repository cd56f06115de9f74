"""Kernel-time breakdown of one training step (torch.profiler / CUPTI).  Debug aid for choosing what to optimise."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sow_b200.trainer import SoWTrainer, TrainConfig  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
tr = SoWTrainer(TrainConfig(batch_size=B), dev)
ids = torch.randint(1, 32000, (B, 256), device=dev)
for i in range(3):
    tr.step(ids)
    if i == 0:
        tr.merge()
torch.cuda.synchronize()
t0 = time.time()
for _ in range(3):
    tr.step(ids)
torch.cuda.synchronize()
print(f"wall per step (no profiler): {(time.time() - t0) / 3 * 1e3:.1f} ms")
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        tr.step(ids)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = {}
for e in ev:
    k = e.name[:90]
    d = tot.setdefault(k, [0.0, 0])
    d[0] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    d[1] += 1
items = sorted(tot.items(), key=lambda kv: -kv[1][0])
total = sum(v[0] for v in tot.values())
print(f"total GPU kernel time per step: {total / 2 / 1e3:.1f} ms over {sum(v[1] for v in tot.values()) // 2} kernels")
for k, (t, n) in items[:40]:
    print(f"{t / 2 / 1e3:8.2f} ms  {n // 2:5d}x  {k}")
