"""TT-QR decomposition / reconstruction / TT-Adam sweep over Llama-7B-shaped weights (BASELINE.json config 5):
4096x4096 and 4096x11008, TT orders 2-3, uniform ranks 8-64.  CUDA-event timed; rates are algorithmic bytes / time
(decompose 4mn + 4r(m+n) per core; fused Adam e_g + 2 e_p + 8 B/element for the new moments).

    python tools/bench_tt.py [--quick]
Prints one JSON line per case and a final summary line.
"""
import json
import os
import sys
from math import ceil

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tn_gradient.optimizer.ttadam import TTAdam  # noqa: E402
from tn_gradient.tt import TensorTrain  # noqa: E402


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    quick = "--quick" in sys.argv
    dev = torch.device("cuda", 0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    shapes = [(4096, 4096), (4096, 11008)]
    orders = [2] if quick else [2, 3]
    rank_list = [8, 64] if quick else [8, 16, 32, 64]
    out = []
    for (M, N) in shapes:
        for order in orders:
            mm, nn = ceil(M ** (1 / order)), ceil(N ** (1 / order))
            P = mm * nn
            for r in rank_list:
                ranks = [1] + [r] * (order - 1) + [1]
                torch.manual_seed(0)
                g32 = torch.randn(M, N, device=dev)
                # --- decomposition / reconstruction of a dense fp32 matrix
                t_dec = timed(lambda: TensorTrain.from_matrix(g32, ranks=ranks, padding=True))
                tt = TensorTrain.from_matrix(g32, ranks=ranks, padding=True)
                t_rec = timed(lambda: tt.to_matrix((M, N)))
                rec = tt.to_matrix((M, N))
                rel = float((rec - g32).norm() / g32.norm())
                # algorithmic bytes of the sweep: every unfolding L_k is read once, Q_k and R_k written once
                dec_bytes, rows, cols, rk = 0, P, P ** (order - 1), 1
                for k in range(order - 1):
                    m_k, n_k = rk * P, cols
                    dec_bytes += 4 * m_k * n_k + 4 * r * (m_k + n_k)
                    rk, cols = r, cols // P
                dec_bytes += 4 * M * N            # reading the source matrix for the pad + interleave pass
                rec_bytes = 4 * M * N + 4 * (P ** order)
                # --- TT-Adam step on a bf16 parameter (steady state: moments already in TT format)
                p = torch.nn.Parameter((torch.randn(M, N, device=dev) * 0.02).bfloat16())
                p.grad = (torch.randn(M, N, device=dev) * 0.01).bfloat16()
                opt = TTAdam([{"params": [p], "ranks": ranks}], lr=1e-3)
                opt.step()
                t_step = timed(lambda: opt.step())
                adam_bytes = M * N * (2 + 2 * 2) + 2 * 4 * (P ** order)      # g, p r/w, new m and v written
                step_bytes = adam_bytes + 2 * (dec_bytes - 4 * M * N)          # + the two decomposition sweeps
                row = {"shape": [M, N], "order": order, "rank": r, "padded": [mm ** order, nn ** order],
                       "decompose_ms": t_dec, "decompose_gbs": dec_bytes / t_dec / 1e6,
                       "reconstruct_ms": t_rec, "reconstruct_gbs": rec_bytes / t_rec / 1e6, "recon_rel_err": rel,
                       "ttadam_step_ms": t_step, "ttadam_step_gbs": step_bytes / t_step / 1e6,
                       "ttadam_frac_of_measured_hbm": step_bytes / t_step / 1e6 / peaks["hbm_gbs"]}
                print(json.dumps(row), flush=True)
                out.append(row)
                del opt, p, tt, g32
                torch.cuda.empty_cache()
    best = max(out, key=lambda d: d["ttadam_frac_of_measured_hbm"])
    worst = min(out, key=lambda d: d["ttadam_frac_of_measured_hbm"])
    print(json.dumps({"summary": "TT sweep", "cases": len(out), "best": best, "worst": worst}))


if __name__ == "__main__":
    main()
