"""Grouped-merge micro-bench: W += s*A.B over every SoW linear of a Llama config in one launch, CUDA-event timed,
reported as algorithmic bytes / time next to MEASURED_PEAKS.json's HBM copy bandwidth.

    python tools/bench_merge.py [llama_350m|llama_7b|roberta] [rank] [iters]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sow_b200 import ops  # noqa: E402

SHAPES = {
    "llama_60m": (512, 1376, 8), "llama_350m": (1024, 2736, 24), "llama_7b": (4096, 11008, 32), "roberta": (768, 3072, 12),
}


def layer_shapes(name):
    if os.environ.get("MERGE_SHAPES"):      # "in,out,count" custom probe
        fin, fout, cnt = [int(v) for v in os.environ["MERGE_SHAPES"].split(",")]
        return [(fin, fout)] * cnt
    h, ff, L = SHAPES[name]
    if name == "roberta":
        per = [(h, h)] * 4 + [(h, ff), (ff, h)]
    else:
        per = [(h, h)] * 4 + [(h, ff)] * 2 + [(ff, h)]
    if os.environ.get("MERGE_SUBSET") == "sq":
        per = [q for q in per if q[0] == q[1]]
    elif os.environ.get("MERGE_SUBSET") == "ragged":
        per = [q for q in per if q[1] == ff]
    elif os.environ.get("MERGE_SUBSET") == "down":
        per = [q for q in per if q[0] == ff]
    return per * L


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "llama_350m"
    r = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    items = []
    nbytes = 0
    for fin, fout in layer_shapes(name):
        W = (torch.randn(fin, fout, device=dev) * 0.02).bfloat16()
        A = (torch.randn(fin, r, device=dev) * 0.05).bfloat16()
        B = (torch.randn(r, fout, device=dev) * 0.05).bfloat16()
        if os.environ.get("MERGE_OOP"):
            items.append((torch.empty_like(W), W, A, B, 1.0))     # out-of-place probe
        else:
            items.append((W, W, A, B, 1.0))
        nbytes += 2 * 2 * fin * fout + 2 * r * (fin + fout)
    # parity of one layer against fp32 torch before timing
    W0 = items[0][0].float().clone()
    ops.merge_grouped(items)
    torch.cuda.synchronize()
    W0 = items[0][1].float().clone() if os.environ.get("MERGE_OOP") else W0
    ref = W0 + items[0][2].float() @ items[0][3].float()
    err = float((items[0][0].float() - ref).norm() / ref.norm())
    for _ in range(2):
        ops.merge_grouped(items)
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.merge_grouped(items)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    # kernel-only time: CUDA events recorded inside the C ABI right around the launch (no host-side table building)
    ops.profile_enable(True)
    for _ in range(iters):
        ops.merge_grouped(items)
    torch.cuda.synchronize()
    k_ms, k_bytes, k_n = ops.profile_read("merge")
    ops.profile_enable(False)
    k_ms /= max(k_n, 1)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    best, med = min(times), sorted(times)[len(times) // 2]
    print(json.dumps({"workload": f"grouped merge {name} r={r} ({len(items)} matrices)", "algorithmic_bytes": nbytes,
                      "ms_best": best, "ms_median": med, "gbs_best": nbytes / best / 1e6, "gbs_median": nbytes / med / 1e6,
                      "frac_of_measured_hbm": nbytes / med / 1e6 / peaks["hbm_gbs"], "rel_err_layer0": err,
                      "kernel_ms": k_ms, "kernel_gbs": nbytes / k_ms / 1e6, "kernel_frac_of_measured_hbm": nbytes / k_ms / 1e6 / peaks["hbm_gbs"]}))


if __name__ == "__main__":
    main()
