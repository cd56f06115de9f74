"""CPU port of the reference's training path -- TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/sow_oracle.py).

The reference (antoine311200/sow) is pure PyTorch and cannot travel to the GPU box, so its CPU path is restated
here with the same torch ops it issues (three ``mm`` + ``mul`` + ``add`` per layer, stack/sum/add merge) and is
timed by ``bench.py`` (``cpu_baseline`` leg and ``--impl reference``).  kind = "port".

Never imported by the product (sow_b200 / tn_gradient).
"""
from __future__ import annotations

import time
from typing import Dict, List

import torch
import torch.nn as nn

LLAMA_SHAPES: Dict[str, Dict[str, int]] = {
    "llama_9m": dict(hidden_size=128, intermediate_size=352, num_hidden_layers=4, num_attention_heads=4),
    "llama_60m": dict(hidden_size=512, intermediate_size=1376, num_hidden_layers=8, num_attention_heads=8),
    "llama_350m": dict(hidden_size=1024, intermediate_size=2736, num_hidden_layers=24, num_attention_heads=16),
}
TARGETS = ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj")


class PortSoWLinear(nn.Module):
    """Eager-torch restatement of SoWLinear (tn_gradient/layer/sow.py:45-178), n_iter = 1, dense accumulation."""

    def __init__(self, fin: int, fout: int, rank: int, scale: float = 1.0, bias: bool = False):
        super().__init__()
        self.fin, self.fout, self.rank, self.scale = fin, fout, rank, scale
        self.W = nn.Parameter(torch.empty(0), requires_grad=False)          # acc_downweight (sow.py:70)
        self.A = nn.Parameter(torch.randn(fin, rank) * 0.02)                # downscale_weights[0]
        self.B = nn.Parameter(torch.randn(rank, fout) * 0.02)               # upscale_weights[0]
        self.bias = nn.Parameter(torch.zeros(fout)) if bias else None

    def forward(self, x):
        low = (x @ self.A @ self.B) * self.scale                            # sow.py:117,119-121
        out = low if self.W.numel() == 0 else (x @ self.W).add_(low)        # sow.py:111-112,119
        if self.bias is not None:
            out = out + self.bias                                           # sow.py:123-124
        return out

    @torch.no_grad()
    def accumulate(self):
        delta = self.scale * torch.sum(torch.stack([self.A.detach() @ self.B.detach()]), dim=0)   # sow.py:131-134
        if self.W.numel() != 0:
            delta = delta + self.W                                          # sow.py:139-140
        self.W = nn.Parameter(delta.contiguous(), requires_grad=False)      # sow.py:152
        self.A.data = torch.zeros_like(self.A).normal_(std=0.02)            # sow.py:174 ("normal": QR init needs CUDA, sow.py:91)
        self.B.data = torch.zeros_like(self.B)                              # sow.py:159,178


def build_port_model(name: str, rank: int, seq_len: int = 256, seed: int = 42) -> nn.Module:
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(vocab_size=32000, max_position_embeddings=max(1024, seq_len), rms_norm_eps=1e-6,
                      hidden_act="silu", initializer_range=0.02, use_cache=False, tie_word_embeddings=False,
                      **LLAMA_SHAPES[name])
    torch.manual_seed(seed)
    model = LlamaForCausalLM(cfg)
    mods = dict(model.named_modules())
    for name_, m in list(model.named_modules()):
        if isinstance(m, nn.Linear) and name_.rsplit(".", 1)[-1] in TARGETS:        # prepare.py:74-83
            new = PortSoWLinear(m.in_features, m.out_features, rank, bias=m.bias is not None)
            parent, child = name_.rsplit(".", 1)
            setattr(mods[parent], child, new)
    return model


def time_cpu_training(model_name: str = "llama_350m", rank: int = 50, batch: int = 2, seq_len: int = 256,
                      steps: int = 2, warmup: int = 1, threads: int = 0, device: str = "cpu",
                      dtype: torch.dtype = torch.float32) -> dict:
    """Training steps of the port in the loop order of scripts/simple_train.py:611-650 (one merge during warm-up so
    the timed steps are the steady state with a dense W).  Returns tokens/s and per-step seconds.

    device="cpu", fp32: the reference's CPU path on the host cores (the reported baseline).  device="cuda", bf16: the
    same eager formulation (three cuBLAS mm + mul + add per layer, stack/sum/add merge, torch.optim.AdamW) on the GPU --
    what the reference itself would run there; an extra baseline leg of bench.py, never part of the product."""
    import os
    n = threads or os.cpu_count() or 1
    if device == "cpu":
        torch.set_num_threads(n)
    model = build_port_model(model_name, rank, seq_len).to(device=device, dtype=dtype)
    sow: List[PortSoWLinear] = [m for m in model.modules() if isinstance(m, PortSoWLinear)]
    factors = [p for m in sow for p in (m.A, m.B)]
    ids = {id(p) for p in factors}
    others = [p for p in model.parameters() if p.requires_grad and id(p) not in ids]
    opt = torch.optim.AdamW([{"params": others, "lr": 1e-2, "weight_decay": 0.0},
                             {"params": factors, "lr": 1e-3, "weight_decay": 0.0}])
    g = torch.Generator().manual_seed(1234)
    times = []
    merge_s = 0.0
    for step in range(warmup + steps):
        ids_ = torch.randint(1, 32000, (batch, seq_len), generator=g).to(device)
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss = model(input_ids=ids_, labels=ids_).loss
        loss.backward()
        if step == 0:
            tm = time.perf_counter()
            for m in sow:
                m.accumulate()
            for p in factors:                                   # reset_optimizer (training_utils.py:257-277)
                st = opt.state[p]
                if st:
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
            merge_s = time.perf_counter() - tm
        opt.step()
        opt.zero_grad()
        if device != "cpu":
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if step >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return {"tokens_per_s": batch * seq_len / sec, "sec_per_step": sec, "merge_sec": merge_s, "threads": n,
            "tokens_per_step": batch * seq_len, "loss": float(loss)}
