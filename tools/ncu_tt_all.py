"""One pass over the TT path for ncu (profile range = cudaProfilerStart/Stop): from_matrix / to_matrix at rank 8 and one
TT-Adam step at ranks 8, 16 and 64 on a 4096 x 4096 bf16 parameter, after warm-up.

    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'tt_|cq_|sum_splits' \
        -o gpurun_out/tt_all python tools/ncu_tt_all.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tn_gradient.optimizer.ttadam import TTAdam  # noqa: E402
from tn_gradient.tt import TensorTrain  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
M = N = 4096
mat = torch.randn(M, N, device=dev)
opts = []
for r in (8, 16, 64):
    p = torch.nn.Parameter((torch.randn(M, N, device=dev) * 0.02).bfloat16())
    p.grad = (torch.randn(M, N, device=dev) * 0.01).bfloat16()
    opt = TTAdam([{"params": [p], "ranks": [1, r, 1]}], lr=1e-3)
    for _ in range(3):
        opt.step()
    opts.append(opt)
tt = TensorTrain.from_matrix(mat, [1, 8, 1])
tt.to_matrix((M, N))
torch.cuda.synchronize()
torch.cuda.profiler.start()
tt = TensorTrain.from_matrix(mat, [1, 8, 1])
tt.to_matrix((M, N))
for opt in opts:
    opt.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
