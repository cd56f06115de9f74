#!/usr/bin/env python
"""Reference arm: the UNMODIFIED reference package (baseline/_ref, installed by baseline/install_ref.sh) driven
through its own public API -- tn_gradient.prepare.prepare_sow / SoWConfig / accumulate and the stock
tn_gradient.layer.sow.SoWLinear forward/backward/merge -- in the loop order of its training scripts
(scripts/simple_train.py:596-650 for pre-training, scripts/run_glue.py:976-1002 for the GLUE-shaped fine-tune), on
synthetic tokens.  BASELINE / TEST INFRASTRUCTURE ONLY: nothing of sow_b200 is imported here, and the product never
imports this file.

Runs as its own process (the reference's package is also called ``tn_gradient``, like the drop-in of this repo, so
the two cannot share an interpreter):

    python baseline/ref_runner.py --model llama_350m --rank 50 --batch 2 --seq 256 --steps 3 --warmup 1 --device cpu
    python baseline/ref_runner.py --model llama_60m --config1                  # BASELINE.json configs[0]
    python baseline/ref_runner.py --model llama_350m --batch 128 --device cuda --dtype bf16

Prints ONE JSON line.  The four packages the reference imports but this image lacks (peft, opt_einsum, termcolor,
galore_torch) are shimmed exactly as in tests/golden/make_golden.py.
"""
import argparse
import json
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, "_ref")

LLAMA_SHAPES = {
    "llama_9m": dict(hidden_size=128, intermediate_size=352, num_hidden_layers=4, num_attention_heads=4),
    "llama_60m": dict(hidden_size=512, intermediate_size=1376, num_hidden_layers=8, num_attention_heads=8),
    "llama_130m": dict(hidden_size=768, intermediate_size=2048, num_hidden_layers=12, num_attention_heads=12),
    "llama_350m": dict(hidden_size=1024, intermediate_size=2736, num_hidden_layers=24, num_attention_heads=16),
    "llama_7b": dict(hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32),
}
LLAMA_TARGETS = ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"]   # simple_train.py:318
ROBERTA_TARGETS = ["query", "key", "value", "output.dense", "intermediate.dense"]               # run_glue.py:572


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "tn_gradient", "layer", "sow.py"))


def install_shims(torch):
    peft = types.ModuleType("peft")

    class PeftConfig:
        def __init__(self, **kwargs):
            for k, v in kwargs.items():
                setattr(self, k, v)

    class PeftModel(torch.nn.Module):
        pass

    peft.PeftConfig, peft.PeftModel = PeftConfig, PeftModel
    sys.modules["peft"] = peft
    tc = types.ModuleType("termcolor")
    tc.colored = lambda s, *a, **k: s
    sys.modules["termcolor"] = tc
    gal = types.ModuleType("galore_torch")
    galp = types.ModuleType("galore_torch.galore_projector")
    galp.GaLoreProjector = object
    gal.galore_projector = galp
    sys.modules["galore_torch"] = gal
    sys.modules["galore_torch.galore_projector"] = galp
    # opt_einsum is only used by tt.py / utils.py helpers that the SoW training path never calls; a stub that raises
    # on use keeps the import chain of tn_gradient.prepare -> utils alive
    oe = types.ModuleType("opt_einsum")
    oec = types.ModuleType("opt_einsum.contract")
    torch.backends.opt_einsum.enabled = False

    def _unused(*a, **k):
        raise RuntimeError("opt_einsum is not installed (not needed on the SoW training path)")

    class ContractExpression:
        pass

    oe.contract = oe.contract_path = oe.contract_expression = _unused
    oec.ContractExpression = ContractExpression
    oe.contract_module = oec
    sys.modules["opt_einsum"] = oe
    sys.modules["opt_einsum.contract"] = oec


def import_reference():
    import torch
    if not available():
        raise RuntimeError(f"{REF_DIR} is missing: run baseline/install_ref.sh in the build container")
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or os.getcwd()) not in (REPO, HERE)]
    sys.path.insert(0, REF_DIR)
    install_shims(torch)
    import tn_gradient.layer.sow as ref_sow
    import tn_gradient.prepare as ref_prepare
    assert os.path.abspath(ref_sow.__file__).startswith(REF_DIR), ref_sow.__file__
    return ref_sow, ref_prepare


def reset_optimizer(optimizer, group_id):
    """scripts/utils/training_utils.py:257-277 (the scripts directory is not part of the installed package)."""
    import torch
    group = optimizer.param_groups[group_id]
    for param in group["params"]:
        state = optimizer.state[param]
        if not state:
            continue
        state["exp_avg"] = torch.zeros_like(param, memory_format=torch.preserve_format)
        state["exp_avg_sq"] = torch.zeros_like(param, memory_format=torch.preserve_format)
        if "step" in state:
            state["step"] = torch.zeros_like(state["step"])


def build_llama(name, seq_len, seed=42):
    import torch
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(vocab_size=32000, max_position_embeddings=max(1024, seq_len), rms_norm_eps=1e-6,
                      hidden_act="silu", initializer_range=0.02, bos_token_id=0, eos_token_id=1, use_cache=False,
                      tie_word_embeddings=False, **LLAMA_SHAPES[name])
    torch.manual_seed(seed)
    return LlamaForCausalLM(cfg)


def build_roberta(seq_len, seed=42):
    import torch
    from transformers import RobertaConfig, RobertaForSequenceClassification
    cfg = RobertaConfig(vocab_size=50265, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                        intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1,
                        attention_probs_dropout_prob=0.1, max_position_embeddings=514, type_vocab_size=1,
                        layer_norm_eps=1e-5, pad_token_id=1, bos_token_id=0, eos_token_id=2, num_labels=2)
    torch.manual_seed(seed)
    return RobertaForSequenceClassification(cfg)


def run(args):
    import torch
    ref_sow, ref_prepare = import_reference()
    device = torch.device(args.device)
    dtype = {"f32": torch.float32, "bf16": torch.bfloat16}[args.dtype]
    threads = os.cpu_count() or 1
    if device.type == "cpu":
        torch.set_num_threads(threads)
    init_method = args.init_method
    if init_method == "normal_QR" and not torch.cuda.is_available():
        init_method = "normal"            # the reference's normal_QR constructor hard-codes .to("cuda") (sow.py:91)

    roberta = args.model.startswith("roberta")
    if roberta:
        model = build_roberta(args.seq)
        for p in model.roberta.parameters():                        # run_glue.py:515-516
            p.requires_grad = False
        targets, decompose, scale = ROBERTA_TARGETS, "keep", args.scale
    else:
        model = build_llama(args.model, args.seq)
        targets, decompose, scale = LLAMA_TARGETS, (None if args.mode == "pretrain" else "keep"), args.scale
        if args.mode == "keep":
            for p in model.parameters():
                p.requires_grad = False
    cfg = ref_prepare.SoWConfig(target_modules=targets, rank=args.rank, scale=scale, device="cpu",
                                init_method=init_method, decompose=decompose)
    model = ref_prepare.prepare_sow(model, cfg)
    sow = [m for m in model.modules() if isinstance(m, ref_sow.SoWLinear)]
    special = [p for m in sow for p in list(m.downscale_weights) + list(m.upscale_weights)]
    for p in special:
        p.requires_grad_(True)
    ids = {id(p) for p in special}
    model = model.to(device=device, dtype=dtype)
    others = [p for p in model.parameters() if p.requires_grad and id(p) not in ids]
    groups = []
    if others:
        groups.append({"params": others, "lr": args.lr, "weight_decay": 0.0})
    sow_gid = len(groups)
    groups.append({"params": special, "lr": args.sow_lr, "weight_decay": 0.0})
    opt = torch.optim.AdamW(groups)
    model.train()
    fwd = torch.compile(model) if args.compile else model          # scripts/finetune.py:486-487

    gen = torch.Generator().manual_seed(1234)
    B, S = args.batch, args.seq

    def batch():
        if roberta:
            ids_ = torch.randint(3, 50265, (B, S), generator=gen)
            return dict(input_ids=ids_.to(device), attention_mask=torch.ones_like(ids_).to(device),
                        labels=torch.randint(0, 2, (B,), generator=gen).to(device))
        ids_ = torch.randint(1, 32000, (B, S), generator=gen).to(device)
        return dict(input_ids=ids_, labels=ids_)

    def sync():
        if device.type == "cuda":
            torch.cuda.synchronize()

    merge_at = set(args.merge_at)
    times, losses, merge_s = [], [], []
    for step in range(args.warmup + args.steps):
        b = batch()
        sync()
        t0 = time.perf_counter()
        loss = fwd(**b).loss
        loss.backward()
        if step in merge_at:                                        # simple_train.py:618-626 / run_glue.py:993-1002
            sync()
            tm = time.perf_counter()
            ref_prepare.accumulate(model)
            reset_optimizer(opt, sow_gid)
            if roberta and not merge_s:
                for m in sow:
                    m.scale = 1.0 / args.rank                        # run_glue.py:996-1001
            sync()
            merge_s.append(time.perf_counter() - tm)
        opt.step()
        opt.zero_grad()
        losses.append(float(loss))
        sync()
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
    timed_merge = sum(1 for s in merge_at if s >= args.warmup)
    sec_total = sum(times)
    out = {
        "impl": "reference", "kind": "reference", "model": args.model, "rank": args.rank, "device": str(device),
        "dtype": args.dtype, "batch": B, "seq": S, "steps": args.steps, "warmup": args.warmup, "threads": threads,
        "tokens_per_step": B * S, "sec_per_step": sec_total / max(len(times), 1),
        "tokens_per_s": B * S * len(times) / sec_total if sec_total > 0 else 0.0,
        "merge_sec": merge_s, "merges_in_timed_steps": timed_merge, "losses": losses, "init_method": init_method,
        "sow_layers": len(sow), "compiled": bool(args.compile),
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama_350m")
    ap.add_argument("--rank", type=int, default=50)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--seq", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--mode", default="pretrain", choices=["pretrain", "keep"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--lr", type=float, default=1e-2)
    ap.add_argument("--sow-lr", type=float, default=1e-3)
    ap.add_argument("--init-method", default="normal_QR")
    ap.add_argument("--merge-at", type=int, nargs="*", default=[0],
                    help="0-based step indices (warm-up included) after whose backward the merge runs")
    ap.add_argument("--compile", action="store_true", help="torch.compile(model) as scripts/finetune.py:486-487 does")
    ap.add_argument("--config1", action="store_true",
                    help="BASELINE.json configs[0]: llama_60m r=50, 16x256, 10 steps incl. one merge, fp32, CPU")
    args = ap.parse_args()
    if args.config1:
        args.model, args.rank, args.batch, args.seq = "llama_60m", 50, 16, 256
        args.steps, args.warmup, args.device, args.dtype, args.merge_at = 10, 0, "cpu", "f32", [5]
    run(args)


if __name__ == "__main__":
    main()
