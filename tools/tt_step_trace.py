"""Timeline of ONE TT-Adam step (kernel start offsets and durations from the torch profiler / CUPTI), debug aid.

    python tools/tt_step_trace.py 4096 4096 8
"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tn_gradient.optimizer.ttadam import TTAdam  # noqa: E402

M, N, r = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (4096, 4096, 8))]
dev = torch.device("cuda", 0)
p = torch.nn.Parameter((torch.randn(M, N, device=dev) * 0.02).bfloat16())
p.grad = (torch.randn(M, N, device=dev) * 0.01).bfloat16()
opt = TTAdam([{"params": [p], "ranks": [1, r, 1]}], lr=1e-3)
for _ in range(5):
    opt.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        opt.step()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type.name == "CUDA"], key=lambda e: e.time_range.start)
n = len(evs) // 3
evs = evs[n:2 * n]                  # the middle step
t0 = evs[0].time_range.start
for e in evs:
    print("%8.1f us  +%6.1f us  %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name[:90]))
print("step span: %.1f us" % (evs[-1].time_range.end - t0))
