"""Short TT-Adam workload for ncu: N steps on one (M, N) bf16 parameter with an order-2 TT of rank r.

    ncu --set full --clock-control none --import-source on -k regex:tt_adam2 --launch-skip 3 -c 1 \
        -o gpurun_out/tt python tools/ncu_tt.py 4096 4096 8
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tn_gradient.optimizer.ttadam import TTAdam  # noqa: E402

M, N, r = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (4096, 4096, 8))]
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 6
dev = torch.device("cuda", 0)
torch.manual_seed(0)
p = torch.nn.Parameter((torch.randn(M, N, device=dev) * 0.02).bfloat16())
p.grad = (torch.randn(M, N, device=dev) * 0.01).bfloat16()
opt = TTAdam([{"params": [p], "ranks": [1, r, 1]}], lr=1e-3)
for _ in range(steps):
    opt.step()
torch.cuda.synchronize()
print("done")
