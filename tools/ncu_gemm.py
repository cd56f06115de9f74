"""Short deterministic workload for ncu: the fused SoW GEMMs of ONE layer shape (default: Llama-350M gate/up projection,
T = 64 x 256 tokens, 1024 -> 2736, r = 50), forward + backward, a few iterations.  Launch order per iteration:
pack, <64> t, <256> y   |   (memset,) K2 (dt + dB), <64,splitK> dA, finalize, <256> dX.  With a 6th argument n > 1 the
layer is a GROUP of n projections sharing x (q/k/v: n = 3, gate/up: n = 2)."""
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
n = int(sys.argv[6]) if len(sys.argv) > 6 else 1
dev = torch.device("cuda", 0)
torch.manual_seed(0)
x = torch.randn(T, fin, device=dev).bfloat16()
dys = [torch.randn(T, fout, device=dev).bfloat16() for _ in range(n)]
Ws = [(torch.randn(fin, fout, device=dev) * 0.02).bfloat16() for _ in range(n)]
As = [(torch.randn(fin, r, device=dev) * 0.05).bfloat16() for _ in range(n)]
Bs = [(torch.randn(r, fout, device=dev) * 0.05).bfloat16() for _ in range(n)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(iters):
    flush.zero_()                                  # inputs > L2 anyway; keep launches cold like inside a step
    ys, A_cat, t_cat = ops.group_fwd(x, [(Ws[i], As[i], Bs[i], None, 1.0) for i in range(n)])
    dx, dAs, dBs, _ = ops.group_bwd(x, A_cat, t_cat, [(Ws[i], Bs[i], dys[i], 1.0, True, True, False) for i in range(n)], True)
torch.cuda.synchronize()
ref = x.float() @ Ws[0].float() + (x.float() @ As[0].float()) @ Bs[0].float()
print("fwd rel err", float((ys[0].float() - ref).norm() / ref.norm()))
