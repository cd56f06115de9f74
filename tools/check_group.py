"""GPU debug / parity sweep of the group kernels (sow_group_fwd / sow_group_bwd: packed factors, shared t_cat, K2 fused
dt + dB pass, split-K dA_cat, multi-segment dX) against a torch fp32 evaluation of the same formulas on the same bf16
inputs.  Prints one line per case; exits non-zero on a failure.  python tools/check_group.py [quick]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sow_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
TOL = 6e-3


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def run_case(T, fin, outs, rs, scales, has_W, has_bias, seed=0, reps=1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    rn = lambda *s, std=1.0: (torch.randn(*s, device=dev, generator=g) * std).bfloat16()
    x = rn(T, fin)
    mem = []
    for out, r, s, hw in zip(outs, rs, scales, has_W):
        W = rn(fin, out, std=0.02) if hw else None
        A = rn(fin, r, std=0.05)
        B = rn(r, out, std=0.05)
        bias = rn(out, std=0.1) if has_bias else None
        dy = rn(T, out)
        mem.append((W, A, B, bias, s, dy))
    ys, A_cat, t_cat = ops.group_fwd(x, [(W, A, B, bias, s) for (W, A, B, bias, s, dy) in mem])
    dx, dAs, dBs, dbs = ops.group_bwd(x, A_cat, t_cat, [(W, B, dy, s, True, True, has_bias) for (W, A, B, bias, s, dy) in mem], True)
    torch.cuda.synchronize()
    xf = x.float()
    errs = {}
    dx_ref = torch.zeros(T, fin, device=dev)
    for i, (W, A, B, bias, s, dy) in enumerate(mem):
        t = (s * (xf @ A.float()))
        y_ref = t.bfloat16().float() @ B.float()
        if W is not None:
            y_ref = y_ref + xf @ W.float()
        if bias is not None:
            y_ref = y_ref + bias.float()
        errs[f"y{i}"] = rel(ys[i], y_ref)
        dyf = dy.float()
        dt = s * (dyf @ B.float().T)
        r = A.shape[1]
        off = sum((rr + 63) // 64 * 64 for rr in rs[:i])
        errs[f"t{i}"] = rel(t_cat[:, off:off + r], t)
        dB_ref = t_cat[:, off:off + r].float().T @ dyf
        dA_ref = xf.T @ dt.bfloat16().float()
        errs[f"dA{i}"] = rel(dAs[i], dA_ref)
        errs[f"dB{i}"] = rel(dBs[i], dB_ref)
        dx_ref += dt.bfloat16().float() @ A.float().T
        if W is not None:
            dx_ref += dyf @ W.float().T
        if bias is not None:
            errs[f"db{i}"] = rel(dbs[i], dyf.sum(0))
    errs["dx"] = rel(dx, dx_ref)
    # bit-reproducibility of the backward
    for _ in range(reps):
        dx2, dAs2, dBs2, _ = ops.group_bwd(x, A_cat, t_cat, [(W, B, dy, s, True, True, False) for (W, A, B, bias, s, dy) in mem], True)
        torch.cuda.synchronize()
        same = torch.equal(dx2, dx) and all(torch.equal(a, b) for a, b in zip(dAs2, dAs)) and all(torch.equal(a, b) for a, b in zip(dBs2, dBs))
        errs["repro"] = 0.0 if same else 1.0
    worst = max(errs.values())
    ok = worst < TOL
    print(("ok  " if ok else "FAIL"), f"T={T} in={fin} outs={outs} rs={rs} W={has_W} bias={has_bias}",
          " ".join(f"{k}={v:.1e}" for k, v in errs.items() if v > 1e-3 or not ok), f"worst={worst:.2e}", flush=True)
    return ok


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    cases = [
        (512, 256, [384], [50], [0.5], [True], False),
        (300, 256, [136], [8], [1.0], [True], True),
        (1000, 512, [512, 512, 512], [50, 50, 50], [1.0, 0.5, 2.0], [True, True, True], False),
        (4096, 1024, [1024, 1024, 1024], [50, 50, 50], [1.0, 1.0, 1.0], [True, True, True], False),
        (4096, 1024, [2736, 2736], [50, 50], [1.0, 1.0], [True, True], False),
        (2048, 2736, [1024], [50], [1.0], [True], False),
        (777, 768, [768, 768, 768], [8, 8, 8], [0.125] * 3, [True] * 3, True),
        (1024, 768, [3072], [8], [0.125], [True], True),
        (640, 512, [512, 1376], [50, 16], [1.0, 1.0], [False, False], False),
        (512, 256, [264], [100], [1.0], [True], False),
        (256, 128, [128, 128], [64, 130], [1.0, 0.25], [True, False], True),
    ]
    if not quick:
        cases += [
            (32768, 1024, [1024, 1024, 1024], [50, 50, 50], [1.0] * 3, [True] * 3, False),
            (32768, 1024, [2736, 2736], [50, 50], [1.0] * 2, [True] * 2, False),
            (32768, 2736, [1024], [50], [1.0], [True], False),
            (8192, 4096, [11008, 11008], [8, 8], [0.125] * 2, [True] * 2, False),
        ]
    ok = True
    for c in cases:
        ok = run_case(*c, reps=2) and ok
    print("ALL OK" if ok else "SOME FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
