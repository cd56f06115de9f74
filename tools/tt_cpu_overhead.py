"""Host-side cost of one TT-Adam step (small matrix so the GPU never limits): optimizer step vs. the ops wrapper vs. the bare
C-ABI call.  Debug aid."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sow_b200 import _lib, ops  # noqa: E402
from tn_gradient.optimizer.ttadam import TTAdam  # noqa: E402

if len(sys.argv) > 3:                      # M N r
    M, N, r = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
else:                                      # M r (square)
    M = N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    r = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
p = torch.nn.Parameter((torch.randn(M, N, device=dev) * 0.02).bfloat16())
p.grad = (torch.randn(M, N, device=dev) * 0.01).bfloat16()
opt = TTAdam([{"params": [p], "ranks": [1, r, 1]}], lr=1e-3)
for _ in range(5):
    opt.step()
torch.cuda.synchronize()


def bench(fn, n=300):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6


print("opt.step            : host %.1f us/step, with final sync %.1f us/step" % bench(opt.step))
st = opt.state[p]
from math import ceil
mm, nn_ = ceil(M ** 0.5), ceil(N ** 0.5)
P = mm * nn_
tm, tv = st["exp_avg"], st["exp_avg_sq"]
cm = (tm.cores[0].reshape(P, -1).contiguous(), tm.cores[1].reshape(-1, P).contiguous())
cv = (tv.cores[0].reshape(P, -1).contiguous(), tv.cores[1].reshape(-1, P).contiguous())
f = lambda: ops.tt_adam2_step(p.data, p.grad, cm, cv, mm, nn_, r, 0.9, 0.999, 1e-8, 1e-3, 0.0, False)
print("ops.tt_adam2_step   : host %.1f us/step, with final sync %.1f us/step" % bench(f))
lib = _lib.load()
Q = torch.empty((2, P, r), dtype=torch.float32, device=dev)
R = torch.empty((2, r, P), dtype=torch.float32, device=dev)
ws = ops.workspace(dev, lib.tt_adam2_workspace_bytes(mm, nn_))
import ctypes
args = [ctypes.c_void_p(t.data_ptr()) for t in (p.data, p.grad, cm[0], cm[1], cv[0], cv[1])]
qa = [ctypes.c_void_p(t.data_ptr()) for t in (Q[0], Q[1], R[0], R[1])]
stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
wsp = ctypes.c_void_p(ws.data_ptr())
g = lambda: lib.tt_adam2_step(*args, r, *qa, M, N, mm, nn_, 0.9, 0.999, 1e-8, 1e-3, 0.0, 0, 0, wsp, ws.numel(), stream)
print("lib.tt_adam2_step   : host %.1f us/step, with final sync %.1f us/step" % bench(g))
