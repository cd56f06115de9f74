"""Kernel-level breakdown of one TTAdam step / from_matrix / to_matrix (torch.profiler, debug aid)."""
import os, sys
import torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tn_gradient.optimizer.ttadam import TTAdam
M, N, order, r = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (4096, 4096, 2, 8))]
ranks = [1] + [r] * (order - 1) + [1]
dev = torch.device("cuda", 0)
p = torch.nn.Parameter((torch.randn(M, N, device=dev) * 0.02).bfloat16())
p.grad = (torch.randn(M, N, device=dev) * 0.01).bfloat16()
opt = TTAdam([{"params": [p], "ranks": ranks}], lr=1e-3)
for _ in range(3):
    opt.step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(10):
    opt.step()
torch.cuda.synchronize()
print("wall per step: %.3f ms" % ((time.perf_counter() - t0) * 100))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        opt.step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 5, e.count / 5) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda x: -x[1])
print("GPU kernel time per step: %.1f us" % sum(x[1] for x in rows))
for k, t, c in rows[:20]:
    print("%9.1f us %5.1fx  %s" % (t, c, k[:110]))

from tn_gradient.tt import TensorTrain
mat = torch.randn(M, N, device=dev)
for name, fn in (("from_matrix", lambda: TensorTrain.from_matrix(mat, list(ranks))),):
    for _ in range(3):
        tt = fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            tt = fn()
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total / 5, e.count / 5) for e in prof.key_averages() if e.device_time_total > 0]
    rows.sort(key=lambda x: -x[1])
    print("%s GPU kernel time per call: %.1f us" % (name, sum(x[1] for x in rows)))
    for k, t, c in rows[:12]:
        print("%9.1f us %5.1fx  %s" % (t, c, k[:110]))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        tt.to_matrix((M, N))
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 5, e.count / 5) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda x: -x[1])
print("to_matrix GPU kernel time per call: %.1f us" % sum(x[1] for x in rows))
for k, t, c in rows[:8]:
    print("%9.1f us %5.1fx  %s" % (t, c, k[:110]))
