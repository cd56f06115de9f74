"""Per-kernel-class timing of one projection GROUP (forward + backward through the C ABI), CUDA events around every
launch (sow_profile_*): python tools/bench_group.py [T].  Shapes: the Llama-350M groups of BASELINE.json config 2."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sow_b200 import ops  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dev = torch.device("cuda", 0)
GROUPS = [("q/k/v", 1024, [1024] * 3, 50), ("o", 1024, [1024], 50), ("gate/up", 1024, [2736] * 2, 50), ("down", 2736, [1024], 50),
          ("roberta qkv", 768, [768] * 3, 8), ("7b gate/up (T/4)", 4096, [11008] * 2, 8)]
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for name, fin, outs, r in GROUPS:
    Tn = T // 4 if fin >= 4096 else T
    torch.manual_seed(0)
    x = torch.randn(Tn, fin, device=dev).bfloat16()
    mem = [((torch.randn(fin, o, device=dev) * 0.02).bfloat16(), (torch.randn(fin, r, device=dev) * 0.05).bfloat16(),
            (torch.randn(r, o, device=dev) * 0.05).bfloat16(), torch.randn(Tn, o, device=dev).bfloat16()) for o in outs]

    def step():
        flush.zero_()
        ys, A_cat, t_cat = ops.group_fwd(x, [(W, A, B, None, 1.0) for W, A, B, dy in mem])
        ops.group_bwd(x, A_cat, t_cat, [(W, B, dy, 1.0, True, True, False) for W, A, B, dy in mem], True)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ops.profile_enable(True)
    iters = 10
    for _ in range(iters):
        step()
    torch.cuda.synchronize()
    row = {"group": name, "T": Tn, "in": fin, "outs": outs, "r": r}
    for k in ("gemm_fwd", "gemm_dx", "gemm_skinny", "gemm_splitk", "gemm_k2"):
        ms, work, n = ops.profile_read(k)
        if n:
            row[k] = {"us_per_launch": round(1e3 * ms / n, 1), "launches_per_iter": n / iters, "tflops": round(work / ms / 1e9, 1)}
    ops.profile_enable(False)
    dyb = sum(Tn * o * 2 for o in outs)
    row["k2_GBps_on_dY"] = round(dyb / len(outs) / (row["gemm_k2"]["us_per_launch"] * 1e-6) / 1e9, 0)
    print(json.dumps(row), flush=True)
    del x, mem
    torch.cuda.empty_cache()
