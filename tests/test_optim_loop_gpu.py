"""GPU: fused multi-tensor AdamW vs the oracle, reset_optimizer rebinding semantics, and the training-loop contract
(scripts/simple_train.py:611-650) replayed against the reference's golden trajectory."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import rel_err
from oracle import sow_oracle as O

pytestmark = pytest.mark.gpu


def reset_optimizer(optimizer, group_id):
    """Behaviour of scripts/utils/training_utils.py:257-277 (rebinding assignment, not in-place zeroing)."""
    group = optimizer.param_groups[group_id]
    for param in group["params"]:
        state = optimizer.state[param]
        state["exp_avg"] = torch.zeros_like(param, memory_format=torch.preserve_format)
        state["exp_avg_sq"] = torch.zeros_like(param, memory_format=torch.preserve_format)
        if group["amsgrad"]:
            state["max_exp_avg_sq"] = torch.zeros_like(param, memory_format=torch.preserve_format)
        if "step" in state:
            state["step"] = torch.zeros_like(state["step"])


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 1e-2)])
def test_fused_adamw_vs_oracle_with_reset(dtype, tol):
    from sow_b200.optim import FusedAdamW
    rng = np.random.default_rng(0)
    shapes = [(1024, 50), (50, 2736), (33,), (70001,)]
    ps_np = [(rng.standard_normal(s) * 0.1).astype(np.float32) for s in shapes]
    if dtype == torch.bfloat16:
        ps_np = [O.bf16_round(p) for p in ps_np]
    params = [nn.Parameter(torch.from_numpy(p).to("cuda", dtype)) for p in ps_np]
    opt = FusedAdamW([{"params": params[:2], "lr": 1e-2, "weight_decay": 0.1},
                      {"params": params[2:], "lr": 3e-3, "weight_decay": 0.0}])
    states = [dict() for _ in shapes]
    cur = [p.astype(np.float64) for p in ps_np]
    for step in range(6):
        grads = [rng.standard_normal(s).astype(np.float32) for s in shapes]
        if dtype == torch.bfloat16:
            grads = [O.bf16_round(gr) for gr in grads]
        for p, gr in zip(params, grads):
            p.grad = torch.from_numpy(gr).to("cuda", dtype)
        if step == 3:
            reset_optimizer(opt, 0)                     # merge step: moments of group 0 rebound to fresh zeros
            states[0], states[1] = {}, {}
        opt.step()
        for i in range(len(shapes)):
            lr, wd = (1e-2, 0.1) if i < 2 else (3e-3, 0.0)
            cur[i] = O.adamw_step(cur[i], grads[i], states[i], lr, weight_decay=wd)
            if dtype == torch.bfloat16:                 # parameters are stored in bf16 between steps
                cur[i] = O.bf16_round(cur[i]).astype(np.float64)
    torch.cuda.synchronize()
    for p, c in zip(params, cur):
        assert rel_err(p.detach().float().cpu().numpy(), c) < tol
    assert float(opt.state[params[0]]["step"]) == 3 and float(opt.state[params[2]]["step"]) == 6


class MLP(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1 = nn.Linear(24, 40, bias=True)
        self.fc2 = nn.Linear(40, 16, bias=False)

    def forward(self, x):
        return self.fc2(torch.tanh(self.fc1(x)))


@pytest.mark.parametrize("mode", ["keep", None])
def test_training_loop_contract_vs_reference(golden_loop, mode):
    """backward -> [accumulate + reset_optimizer at step 3] -> optimizer.step, on the reference's own initial
    weights and batches.  Up to the merge the trajectory must follow the reference within bf16 tolerance; the
    merged W (deterministic) must match; after the merge A is a fresh random draw, so only invariants hold."""
    from tn_gradient.prepare import SoWConfig, accumulate, prepare_sow
    g = golden_loop
    tag = f"loop/{mode}"
    model = MLP()
    with torch.no_grad():
        model.fc1.weight.copy_(torch.from_numpy(g[f"{tag}/fc1_weight"]))
        model.fc1.bias.copy_(torch.from_numpy(g[f"{tag}/fc1_bias"]))
        model.fc2.weight.copy_(torch.from_numpy(g[f"{tag}/fc2_weight"]))
    prepare_sow(model, SoWConfig(target_modules=["fc1", "fc2"], rank=4, scale=0.5, device="cpu",
                                 init_method="normal", decompose=mode))
    factors = []
    for mod in [model.fc1, model.fc2]:
        factors += list(mod.downscale_weights) + list(mod.upscale_weights)
    with torch.no_grad():
        for i, f in enumerate(factors):
            f.copy_(torch.from_numpy(g[f"{tag}/factor{i}_init"]))
    model = model.to("cuda")
    others = [p for p in model.parameters() if p.requires_grad and all(p is not f for f in factors)]
    opt = torch.optim.AdamW([{"params": others, "lr": 1e-2, "weight_decay": 0.0},
                             {"params": factors, "lr": 1e-2, "weight_decay": 0.0}])
    losses = []
    ids = [id(f) for f in factors]
    for step in range(1, 7):
        x = torch.from_numpy(g[f"{tag}/x"][step - 1]).cuda()
        y = torch.from_numpy(g[f"{tag}/y"][step - 1]).cuda()
        loss = ((model(x) - y) ** 2).mean()
        loss.backward()
        if step == 3:
            accumulate(model)
            reset_optimizer(opt, 1)
            assert all(float(st["exp_avg"].abs().max()) == 0 for st in (opt.state[f] for f in factors))
            assert float(model.fc1.upscale_weights[0].abs().max()) == 0.0
        opt.step()
        opt.zero_grad()
        losses.append(float(loss))
    ref = g[f"{tag}/losses"]
    assert np.allclose(losses[:3], ref[:3], rtol=2e-2), (losses, ref)
    assert np.all(np.isfinite(losses)) and abs(losses[5] - ref[5]) < 0.2 * ref[5]
    assert [id(f) for f in factors] == ids                                # Parameter identity survives the merge
    assert rel_err(model.fc1.acc_downweight.float().cpu().numpy(), g[f"{tag}/final_W_fc1"]) < 2e-2
    assert rel_err(model.fc2.acc_downweight.float().cpu().numpy(), g[f"{tag}/final_W_fc2"]) < 2e-2
    assert sorted(model.state_dict().keys()) == sorted(g[f"{tag}/state_dict_keys"].tolist())
