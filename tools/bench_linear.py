"""SoW linear fwd / fwd+bwd micro-bench: the kernel-backed layer vs the reference's own eager formulation on the same
GPU (three cuBLAS mm + mul + add through autograd, tn_gradient/layer/sow.py:107-126) -- the "practical bar" of
BASELINE.md section 5.  bf16, CUDA-event timed, inputs larger than L2.

    python tools/bench_linear.py [T]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tn_gradient.layer.sow import SoWLinear  # noqa: E402

SHAPES = [  # name, in, out, r, scale
    ("llama_350m q/k/v/o", 1024, 1024, 50, 1.0), ("llama_350m gate/up", 1024, 2736, 50, 1.0),
    ("llama_350m down", 2736, 1024, 50, 1.0), ("roberta intermediate", 768, 3072, 8, 0.125),
    ("llama_7b q/k/v/o", 4096, 4096, 8, 0.125), ("llama_7b gate/up", 4096, 11008, 8, 0.125),
]


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    dev = torch.device("cuda", 0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops_sustained": 1400.0}
    for name, fin, fout, r, scale in SHAPES:
        Tn = T if fin * fout < 3e7 else T // 4          # keep the 7B shapes to a few GB of activations
        torch.manual_seed(0)
        layer = SoWLinear(fin, fout, bias=False, rank=r, scale=scale, init_method="normal", dtype=torch.bfloat16, device=dev)
        layer.acc_downweight = torch.nn.Parameter((torch.randn(fin, fout, device=dev) * 0.02).bfloat16(), requires_grad=False)
        with torch.no_grad():
            layer.upscale_weights[0].normal_(0, 0.02)
        W, A, B = layer.acc_downweight, layer.downscale_weights[0], layer.upscale_weights[0]
        x = torch.randn(Tn, fin, device=dev, dtype=torch.bfloat16, requires_grad=True)
        dy = torch.randn(Tn, fout, device=dev, dtype=torch.bfloat16)

        def ref_fwd():
            return x @ W + ((x @ A) @ B) * scale          # sow.py:112,117,119-121

        def ours_fb():
            x.grad = A.grad = B.grad = None
            layer(x).backward(dy)

        def ref_fb():
            x.grad = A.grad = B.grad = None
            ref_fwd().backward(dy)

        with torch.no_grad():
            t_of, t_rf = timed(lambda: layer(x)), timed(ref_fwd)
        t_ofb, t_rfb = timed(ours_fb), timed(ref_fb)
        flops_f = 2.0 * Tn * fin * fout + 2.0 * Tn * r * (fin + fout)
        flops_fb = 4.0 * Tn * fin * fout + 6.0 * Tn * r * (fin + fout)
        print(json.dumps({"shape": name, "T": Tn, "in": fin, "out": fout, "r": r,
                          "fwd_ms": t_of, "fwd_ms_eager_reference": t_rf, "fwd_speedup": t_rf / t_of,
                          "fwd_tflops": flops_f / t_of / 1e9,
                          "fwd_bwd_ms": t_ofb, "fwd_bwd_ms_eager_reference": t_rfb, "fwd_bwd_speedup": t_rfb / t_ofb,
                          "fwd_bwd_tflops": flops_fb / t_ofb / 1e9,
                          "fwd_bwd_frac_of_sustained_cublas": flops_fb / t_ofb / 1e9 / peaks["bf16_tflops_sustained"]}), flush=True)
        del layer, x, dy
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
