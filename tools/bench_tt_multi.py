"""TT-Adam over SEVERAL independent parameters (what an optimizer step on a model is): time per parameter with the
tensor-train updates on one stream and alternating between side streams (SOWB_TT_STREAMS), and a bit-equality check of
the two.

    python tools/bench_tt_multi.py [n_params] [M] [N] [rank]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tn_gradient.optimizer.ttadam import TTAdam  # noqa: E402

n_params = int(sys.argv[1]) if len(sys.argv) > 1 else 8
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
N = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
r = int(sys.argv[4]) if len(sys.argv) > 4 else 8
dev = torch.device("cuda", 0)


def run(streams):
    os.environ["SOWB_TT_STREAMS"] = str(streams)
    torch.manual_seed(0)
    ps = [torch.nn.Parameter((torch.randn(M, N, device=dev) * 0.02).bfloat16()) for _ in range(n_params)]
    for p in ps:
        p.grad = (torch.randn(M, N, device=dev) * 0.01).bfloat16()
    opt = TTAdam([{"params": ps, "ranks": [1, r, 1]}], lr=1e-3)
    for _ in range(3):
        opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 / n_params * 1e3, [p.detach().clone() for p in ps]


res = {k: run(k) for k in (1, 2, 3, 4)}
same = all(torch.equal(a, b) for k in (2, 3, 4) for a, b in zip(res[1][1], res[k][1]))
print(f"{n_params} x ({M} x {N}) rank {r}: " + ", ".join(f"{k} stream{'s' if k > 1 else ''} {res[k][0]:.1f} us" for k in res)
      + f" per parameter; results bit-identical: {same}")
