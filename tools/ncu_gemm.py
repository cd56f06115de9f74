"""Short deterministic workload for ncu: the fused SoW GEMMs of ONE layer shape (default: Llama-350M gate/up projection,
T = 64 x 256 tokens, 1024 -> 2736, r = 50), forward + backward, a few iterations.  Launch order per iteration:
pack, <64> t, <256> y   |   memset, <64> dt, <64,splitK> dB, <64,splitK> dA, finalize, pack, <256> dX."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sow_b200 import ops  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
fin = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
fout = int(sys.argv[3]) if len(sys.argv) > 3 else 2736
r = int(sys.argv[4]) if len(sys.argv) > 4 else 50
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 4
dev = torch.device("cuda", 0)
torch.manual_seed(0)
x = torch.randn(T, fin, device=dev).bfloat16()
dy = torch.randn(T, fout, device=dev).bfloat16()
W = (torch.randn(fin, fout, device=dev) * 0.02).bfloat16()
A = (torch.randn(fin, r, device=dev) * 0.05).bfloat16()
B = (torch.randn(r, fout, device=dev) * 0.05).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(iters):
    flush.zero_()                                  # inputs > L2 anyway; keep launches cold like inside a step
    y, t = ops.linear_fwd(x, W, A, B, None, 1.0)
    dt, dA, dB, _ = ops.linear_bwd_factors(dy, x, t, B, 1.0, False, fin)
    dx = ops.linear_bwd_dx(dy, dt, W, A)
torch.cuda.synchronize()
ref = x.float() @ W.float() + (x.float() @ A.float()) @ B.float()
print("fwd rel err", float((y.float() - ref).norm() / ref.norm()))
