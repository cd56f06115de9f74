"""BASELINE.json configs[0]: llama_60m, SoW rank 50, synthetic tokens batch 16 x seq 256, 10 steps incl. one merge (after the
backward of step 6: update_step 5, sow_accumulation 5), fp32.

The reference runs it on the host CPU (baseline/ref_runner.py --config1: the unmodified package from baseline/_ref); this
build runs the same loop on the GPU through the fp32-faithful path (fp32 parameters, bf16x3 tensor-core products, exact
fp32 merge).  Prints one JSON line with both timings and both loss trajectories (same token stream, different random
initial weights: the reference's init RNG is its own).
    python tools/config1.py [--no-reference]
"""
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sow_b200.trainer import SoWTrainer, TrainConfig  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    cfg = TrainConfig(model="llama_60m", rank=50, seq_len=256, batch_size=16, sow_accumulation=5, dtype=torch.float32,
                      init_method="normal_QR")
    tr = SoWTrainer(cfg, dev)
    gen = torch.Generator().manual_seed(1234)
    times, losses = [], []
    for step in range(10):
        ids = torch.randint(1, 32000, (16, 256), generator=gen).to(dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss = tr.step(ids)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        losses.append(float(loss))
    ours = {"ms_per_step_all": [round(t * 1e3, 2) for t in times], "ms_per_step_median_excl_first": round(sorted(times[1:])[4] * 1e3, 2),
            "tokens_per_s": 16 * 256 / sorted(times[1:])[4], "merges": tr.merges, "losses": losses,
            "dtype": "fp32 parameters / activations, bf16x3 tensor-core products, exact fp32 merge"}
    ref = None
    if "--no-reference" not in sys.argv:
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
        res = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "ref_runner.py"), "--config1"], capture_output=True,
                             text=True, env=env, cwd=os.path.join(ROOT, "baseline"))
        for ln in reversed(res.stdout.strip().splitlines()):
            if ln.startswith("{"):
                r = json.loads(ln)
                ref = {"sec_per_step": r["sec_per_step"], "tokens_per_s": r["tokens_per_s"], "threads": r["threads"],
                       "merge_sec": r["merge_sec"], "losses": r["losses"], "kind": "reference (baseline/_ref) on the host CPU, fp32"}
                break
    print(json.dumps({"config": "BASELINE.json configs[0]: llama_60m r=50, 16x256 tokens, 10 steps incl. one merge, fp32",
                      "this_build_gpu": ours, "reference_cpu": ref,
                      "speedup": None if ref is None else ours["tokens_per_s"] / ref["tokens_per_s"]}))


if __name__ == "__main__":
    main()
