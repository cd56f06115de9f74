"""GPU parity: SoWLinear forward/backward through the public tn_gradient API (-> C ABI -> tcgen05 kernels) against
the CPU oracle on the same inputs and against outputs of the unmodified reference (tests/golden/linear.npz).

Tolerance (BASELINE.json north_star): rel. error <= 1e-2 vs an fp32/fp64 reference for bf16 compute."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import sow_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-2


def _layer_from_arrays(fin, fout, r, n_iter, scale, As, Bs, W, bias, dtype):
    from tn_gradient.layer.sow import SoWLinear
    layer = SoWLinear(fin, fout, bias=bias is not None, rank=r, n_iter=n_iter, scale=scale, init_method="normal",
                      dtype=torch.float32)
    with torch.no_grad():
        for i in range(n_iter):
            layer.downscale_weights[i].copy_(torch.from_numpy(As[i]))
            layer.upscale_weights[i].copy_(torch.from_numpy(Bs[i]))
        if bias is not None:
            layer.bias.copy_(torch.from_numpy(bias))
    if W is not None:
        layer.acc_downweight = torch.nn.Parameter(torch.from_numpy(W).clone(), requires_grad=False)
    layer.virtual_rank = min(fin, fout)
    return layer.to("cuda", dtype)


@pytest.mark.parametrize("case", ["f32_w_bias", "f32_now", "f32_niter2", "bf16_w", "bf16_now_bias"])
def test_forward_backward_vs_reference_golden_and_oracle(golden_linear, case):
    g = golden_linear
    p = f"linear/{case}/"
    n_iter = int(g[p + "n_iter"])
    As = [g[p + f"A{i}"] for i in range(n_iter)]
    Bs = [g[p + f"B{i}"] for i in range(n_iter)]
    W = g[p + "W"] if (p + "W") in g else None
    bias = g[p + "bias"] if (p + "bias") in g else None
    scale = float(g[p + "scale"])
    dtype = torch.bfloat16 if case.startswith("bf16") else torch.float32
    fin, r = As[0].shape
    fout = Bs[0].shape[1]
    layer = _layer_from_arrays(fin, fout, r, n_iter, scale, As, Bs, W, bias, dtype)
    x = torch.from_numpy(g[p + "x"]).to("cuda", dtype).requires_grad_(True)
    dy = torch.from_numpy(g[p + "dy"]).to("cuda", dtype)
    y = layer(x)
    assert y.shape == dy.shape and y.dtype == dtype
    y.backward(dy)
    torch.cuda.synchronize()

    # oracle on the same inputs (fp64 truth)
    y_o = O.sow_linear_forward(g[p + "x"], W, As, Bs, bias, scale)
    dx_o, dA_o, dB_o, db_o = O.sow_linear_backward(g[p + "dy"], g[p + "x"], W, As, Bs, scale)
    got = {"y": y, "dx": x.grad}
    want_o = {"y": y_o, "dx": dx_o}
    for i in range(n_iter):
        got[f"dA{i}"], got[f"dB{i}"] = layer.downscale_weights[i].grad, layer.upscale_weights[i].grad
        want_o[f"dA{i}"], want_o[f"dB{i}"] = dA_o[i], dB_o[i]
    if bias is not None:
        got["dbias"], want_o["dbias"] = layer.bias.grad, db_o
    for k in got:
        a = got[k].detach().float().cpu().numpy()
        assert rel_err(a, want_o[k]) < TOL, (k, "oracle", rel_err(a, want_o[k]))
        assert rel_err(a, g[p + k]) < 1.5 * TOL, (k, "reference golden", rel_err(a, g[p + k]))


SHAPES = [
    # T, in, out, r, scale, W, bias   (BASELINE.json configs)
    (4096, 1024, 1024, 50, 1.0, True, False),     # llama_350m q/k/v/o
    (4096, 1024, 2736, 50, 1.0, True, False),     # llama_350m gate/up
    (4096, 2736, 1024, 50, 1.0, True, False),     # llama_350m down
    (4096, 512, 1376, 50, 1.0, False, False),     # llama_60m, pre-merge phase (no W)
    (8192, 768, 3072, 8, 0.125, True, True),      # roberta-base intermediate.dense
    (2048, 4096, 11008, 8, 0.125, True, False),   # llama_7b gate/up
    (300, 264, 136, 50, 1.0, True, True),         # ragged everything
    (1, 64, 64, 3, 2.0, True, False),             # single token
    (130, 128, 128, 100, 1.0, True, False),       # rank > 64 (two tail k-blocks)
]


@pytest.mark.parametrize("T,fin,fout,r,scale,has_W,has_bias", SHAPES)
def test_forward_backward_vs_oracle_at_baseline_shapes(T, fin, fout, r, scale, has_W, has_bias):
    rng = np.random.default_rng(1234)
    x = O.bf16_round(rng.standard_normal((T, fin), dtype=np.float32))
    A = O.bf16_round(rng.standard_normal((fin, r), dtype=np.float32) * 0.05)
    B = O.bf16_round(rng.standard_normal((r, fout), dtype=np.float32) * 0.05)
    W = O.bf16_round(rng.standard_normal((fin, fout), dtype=np.float32) * 0.02) if has_W else None
    bias = O.bf16_round(rng.standard_normal(fout, dtype=np.float32)) if has_bias else None
    dy = O.bf16_round(rng.standard_normal((T, fout), dtype=np.float32))
    layer = _layer_from_arrays(fin, fout, r, 1, scale, [A], [B], W, bias, torch.bfloat16)
    xt = torch.from_numpy(x).to("cuda", torch.bfloat16).requires_grad_(True)
    y = layer(xt)
    y.backward(torch.from_numpy(dy).to("cuda", torch.bfloat16))
    torch.cuda.synchronize()
    # fp32 BLAS oracle is enough here (error budget 1e-2) and keeps the CPU side to seconds
    y_o = O.sow_linear_forward(x, W, [A], [B], bias, scale, dtype=np.float32)
    dx_o, dA_o, dB_o, db_o = O.sow_linear_backward(dy, x, W, [A], [B], scale, dtype=np.float32)
    assert rel_err(y.detach().float().cpu().numpy(), y_o) < TOL
    assert rel_err(xt.grad.float().cpu().numpy(), dx_o) < TOL
    assert rel_err(layer.downscale_weights[0].grad.float().cpu().numpy(), dA_o[0]) < TOL
    assert rel_err(layer.upscale_weights[0].grad.float().cpu().numpy(), dB_o[0]) < TOL
    if has_bias:
        assert rel_err(layer.bias.grad.float().cpu().numpy(), db_o) < TOL


def test_linearity_and_scale_property_at_full_size():
    """Size-independent properties at a BASELINE shape: y is linear in x and affine in `scale`."""
    from tn_gradient.layer.sow import SoWLinear
    torch.manual_seed(0)
    layer = SoWLinear(1024, 2736, bias=False, rank=50, init_method="normal", dtype=torch.bfloat16, device="cuda")
    layer.acc_downweight = torch.nn.Parameter((torch.randn(1024, 2736, device="cuda") * 0.02).bfloat16(), requires_grad=False)
    with torch.no_grad():
        layer.upscale_weights[0].normal_(0, 0.05)
    x1 = torch.randn(16, 256, 1024, device="cuda").bfloat16()
    x2 = torch.randn(16, 256, 1024, device="cuda").bfloat16()
    with torch.no_grad():
        y1, y2, y12 = layer(x1).float(), layer(x2).float(), layer(x1 + x2).float()
        e_lin = float((y12 - (y1 + y2)).norm() / y12.norm())
        layer.scale = 0.0
        y_base = layer(x1).float()
        layer.scale = 2.0
        y_two = layer(x1).float()
        e_aff = float(((y_two - y_base) - 2.0 * (y1 - y_base)).norm() / (y_two - y_base).norm())
    assert e_lin < 2e-2 and e_aff < 2e-2, (e_lin, e_aff)


def test_frozen_factors_and_no_input_grad():
    from tn_gradient.layer.sow import SoWLinear
    layer = SoWLinear(128, 128, bias=False, rank=8, init_method="normal", dtype=torch.bfloat16, device="cuda")
    layer.upscale_weights[0].requires_grad_(False)
    x = torch.randn(64, 128, device="cuda", dtype=torch.bfloat16)      # x does not require grad
    layer(x).sum().backward()
    assert layer.downscale_weights[0].grad is not None and layer.upscale_weights[0].grad is None


def test_activation_checkpointing_reentrant_and_not():
    from torch.utils.checkpoint import checkpoint
    from tn_gradient.layer.sow import SoWLinear
    torch.manual_seed(0)
    layer = SoWLinear(256, 256, bias=True, rank=8, init_method="normal", dtype=torch.bfloat16, device="cuda")
    with torch.no_grad():
        layer.upscale_weights[0].normal_(0, 0.05)
        layer.bias.normal_()
    x = torch.randn(4, 32, 256, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    layer(x).float().pow(2).sum().backward()
    ref = [x.grad.clone(), layer.downscale_weights[0].grad.clone(), layer.bias.grad.clone()]
    for reentrant in (False, True):
        x.grad = None
        layer.zero_grad()
        checkpoint(layer, x, use_reentrant=reentrant).float().pow(2).sum().backward()
        for a, b in zip([x.grad, layer.downscale_weights[0].grad, layer.bias.grad], ref):
            assert torch.equal(a, b)


def test_unsupported_feature_size_fails_loudly():
    from sow_b200 import SowB200Error
    from tn_gradient.layer.sow import SoWLinear
    layer = SoWLinear(100, 5461 - 5461 % 2 + 1, bias=False, rank=4, init_method="normal", dtype=torch.bfloat16, device="cuda")
    with pytest.raises(SowB200Error, match="multiple of 8"):
        layer(torch.randn(4, 100, device="cuda", dtype=torch.bfloat16))


def test_model_with_sow_layers_runs_under_torch_compile():
    """scripts/finetune.py:486-487 wraps the model in torch.compile: the layers are registered torch.library custom ops
    (sow_b200::linear_fwd / linear_bwd with fake implementations and an autograd formula), so dynamo traces through them
    with ZERO graph breaks, and the compiled module gives the same numbers as the eager one."""
    import torch._dynamo
    import torch.nn as nn
    from tn_gradient.prepare import SoWConfig, prepare_sow
    torch.manual_seed(0)
    m = nn.Sequential(nn.Linear(256, 512), nn.GELU(), nn.Linear(512, 256))
    m = prepare_sow(m, SoWConfig(target_modules=["0", "2"], rank=8, device="cuda", init_method="normal",
                                 decompose="keep")).to("cuda", torch.bfloat16)
    x = torch.randn(64, 256, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    y0 = m(x)
    y0.float().pow(2).mean().backward()
    g0 = x.grad.clone()
    gp0 = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    x.grad = None
    m.zero_grad()
    torch._dynamo.reset()
    ex = torch._dynamo.explain(m)(x)
    assert ex.graph_break_count == 0, ex.break_reasons
    assert ex.graph_count == 1
    y1 = torch.compile(m, fullgraph=True)(x)
    y1.float().pow(2).mean().backward()
    assert torch.equal(y0, y1) and torch.equal(g0, x.grad)
    for n, p in m.named_parameters():
        if n in gp0:
            assert torch.equal(gp0[n], p.grad), n


def test_fp32_module_under_torch_compile_uses_the_fp32_path():
    import torch._dynamo
    import torch.nn as nn
    from tn_gradient.prepare import SoWConfig, accumulate, prepare_sow
    torch.manual_seed(0)
    m = nn.Sequential(nn.Linear(256, 512), nn.Tanh(), nn.Linear(512, 256))
    m = prepare_sow(m, SoWConfig(target_modules=["0", "2"], rank=8, device="cuda", init_method="normal", decompose="keep")).to("cuda")
    x = torch.randn(64, 256, device="cuda")
    torch._dynamo.reset()
    cm = torch.compile(m, fullgraph=True)
    with torch.no_grad():
        y_e, y_c = m(x), cm(x)
        assert y_c.dtype == torch.float32 and torch.equal(y_e, y_c)
        for mod in (m[0], m[2]):
            mod.upscale_weights[0].normal_(0, 0.05)
        accumulate(m)                                  # in-place fp32 merge: the cached bf16 pieces of W must be dropped
        y_e2, y_c2 = m(x), cm(x)
        assert torch.equal(y_e2, y_c2) and not torch.equal(y_e, y_e2)


def test_empty_batch_returns_empty_output_and_zero_grads():
    from tn_gradient.layer.sow import SoWLinear
    layer = SoWLinear(64, 128, bias=True, rank=8, init_method="normal", dtype=torch.bfloat16, device="cuda")
    x = torch.empty(0, 7, 64, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    y = layer(x)
    assert y.shape == (0, 7, 128)
    y.sum().backward()
    assert x.grad.shape == x.shape
    assert float(layer.downscale_weights[0].grad.abs().max()) == 0.0 and float(layer.bias.grad.abs().max()) == 0.0


def test_backward_is_bit_reproducible():
    """dA / dB contract over T with split-K; the partial sums are reduced in a fixed order (no fp32 atomics), so two
    backward passes over the same inputs give identical bits -- as the reference's cuBLAS path does."""
    from tn_gradient.layer.sow import SoWLinear
    torch.manual_seed(0)
    layer = SoWLinear(1024, 2736, bias=True, rank=50, init_method="normal", dtype=torch.bfloat16, device="cuda")
    layer.acc_downweight = torch.nn.Parameter((torch.randn(1024, 2736, device="cuda") * 0.02).bfloat16(), requires_grad=False)
    with torch.no_grad():
        layer.upscale_weights[0].normal_(0, 0.05)
    x = torch.randn(8192, 1024, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    dy = torch.randn(8192, 2736, device="cuda", dtype=torch.bfloat16)
    runs = []
    for _ in range(3):
        x.grad = None
        layer.zero_grad()
        layer(x).backward(dy)
        runs.append([x.grad.clone(), layer.downscale_weights[0].grad.clone(), layer.upscale_weights[0].grad.clone(),
                     layer.bias.grad.clone()])
    for other in runs[1:]:
        for a, b in zip(runs[0], other):
            assert torch.equal(a, b)


def test_backward_as_first_cuda_work_of_the_autograd_thread():
    """The first node autograd's worker thread runs may be ours: no CUDA call has bound a context to that thread yet, and the
    TMA descriptor encode (a driver call) needs one.  Fresh shapes defeat the descriptor cache."""
    import threading
    from tn_gradient.layer.sow import SoWLinear
    out = {}

    def body():
        layer = SoWLinear(328, 456, bias=False, rank=8, init_method="normal", dtype=torch.bfloat16, device="cuda")
        x = torch.randn(77, 328, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        dy = torch.randn(77, 456, device="cuda", dtype=torch.bfloat16)
        y = layer(x)
        y.backward(dy)                     # backward thread: first work is sow_group_bwd on new pointers
        torch.cuda.synchronize()
        out["ok"] = bool(torch.isfinite(x.grad.float()).all())

    th = threading.Thread(target=body)
    th.start()
    th.join()
    assert out.get("ok") is True


# ---------------------------------------------------------------------------------------------------------
# projections that share their input (q/k/v, gate/up): one autograd node per group
# ---------------------------------------------------------------------------------------------------------

class _Block(torch.nn.Module):
    """q/k/v + gate/up call pattern of a transformer block: siblings are called with the identical tensor."""

    def __init__(self, h, ff, bias=False):
        super().__init__()
        self.q_proj = torch.nn.Linear(h, h, bias=bias)
        self.k_proj = torch.nn.Linear(h, h // 2, bias=bias)          # GQA-style narrower k/v
        self.v_proj = torch.nn.Linear(h, h // 2, bias=bias)
        self.gate_proj = torch.nn.Linear(h, ff, bias=bias)
        self.up_proj = torch.nn.Linear(h, ff, bias=bias)
        self.down_proj = torch.nn.Linear(ff, h, bias=bias)

    def forward(self, x):
        q, k, v = self.q_proj(x), self.k_proj(x), self.v_proj(x)
        a = q + torch.cat([k, v], dim=-1)
        return self.down_proj(torch.nn.functional.silu(self.gate_proj(a)) * self.up_proj(a))


def _mk_block(h, ff, r, fuse, decompose="keep", bias=False, seed=0):
    from tn_gradient.prepare import SoWConfig, prepare_sow
    torch.manual_seed(seed)
    blk = _Block(h, ff, bias=bias)
    cfg = SoWConfig(target_modules=["q_proj", "k_proj", "v_proj", "gate_proj", "up_proj", "down_proj"], rank=r,
                    device="cuda", init_method="normal", decompose=decompose, fuse_shared_input=fuse)
    blk = prepare_sow(blk, cfg).to("cuda", torch.bfloat16)
    torch.manual_seed(seed + 1)
    with torch.no_grad():
        for m in blk.modules():
            if hasattr(m, "upscale_weights"):
                m.upscale_weights[0].normal_(0, 0.05)
    return blk


@pytest.mark.parametrize("decompose", ["keep", None])
def test_shared_input_groups_match_ungrouped_execution(decompose):
    """Grouped (one t_cat / dA_cat / dX launch per q/k/v and gate/up) vs one node per projection: same function, same
    gradients up to the bf16 rounding of the differently-ordered fp32 sums."""
    from sow_b200.layer import SharedInputGroup
    h, ff, r = 512, 1376, 50
    fused = _mk_block(h, ff, r, True, decompose)
    plain = _mk_block(h, ff, r, False, decompose)
    assert isinstance(fused.q_proj._group, SharedInputGroup) and fused.q_proj._group is fused.v_proj._group
    assert fused.gate_proj._group is fused.up_proj._group and fused.down_proj._group is None
    assert plain.q_proj._group is None
    plain.load_state_dict(fused.state_dict())
    x = torch.randn(4, 130, h, device="cuda", dtype=torch.bfloat16)
    outs = []
    for blk in (fused, plain):
        xi = x.clone().requires_grad_(True)
        y = blk(xi)
        y.float().pow(2).mean().backward()
        outs.append((y, xi.grad, {n: p.grad for n, p in blk.named_parameters() if p.grad is not None}))
    torch.cuda.synchronize()
    (y0, gx0, g0), (y1, gx1, g1) = outs
    assert fused.q_proj._group.enabled and fused.q_proj._group.misses == 0 and fused.q_proj._group.cache is None
    assert float((y0.float() - y1.float()).norm() / y1.float().norm()) < 5e-3
    assert float((gx0.float() - gx1.float()).norm() / gx1.float().norm()) < 1e-2
    assert set(g0) == set(g1) and len(g0) == 12
    for k in g0:
        assert float((g0[k].float() - g1[k].float()).norm() / (g1[k].float().norm() + 1e-20)) < 1e-2, k


def test_shared_input_group_vs_oracle_with_bias_and_distinct_scales():
    from tn_gradient.prepare import SoWConfig, prepare_sow

    class Attn(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.query, self.key, self.value = (torch.nn.Linear(768, 768) for _ in range(3))

        def forward(self, x):
            return self.query(x), self.key(x), self.value(x)

    torch.manual_seed(3)
    att = prepare_sow(Attn(), SoWConfig(target_modules=["query", "key", "value"], rank=8, device="cuda",
                                        init_method="normal", decompose="keep")).to("cuda", torch.bfloat16)
    mods = [att.query, att.key, att.value]
    for i, m in enumerate(mods):
        m.scale = [0.125, 1.0, 0.5][i]
        with torch.no_grad():
            m.upscale_weights[0].normal_(0, 0.05)
            m.bias.normal_(0, 0.1)
    x = torch.randn(1000, 768, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    dys = [torch.randn(1000, 768, device="cuda", dtype=torch.bfloat16) for _ in range(3)]
    ys = att(x)
    torch.autograd.backward(ys, dys)
    torch.cuda.synchronize()
    xn = x.detach().float().cpu().numpy()
    dx_o = 0
    for m, y, dy in zip(mods, ys, dys):
        W, A, B = (t.detach().float().cpu().numpy() for t in (m.acc_downweight, m.downscale_weights[0], m.upscale_weights[0]))
        b = m.bias.detach().float().cpu().numpy()
        dyn = dy.float().cpu().numpy()
        y_o = O.sow_linear_forward(xn, W, [A], [B], b, m.scale, dtype=np.float32)
        dxi, dA_o, dB_o, db_o = O.sow_linear_backward(dyn, xn, W, [A], [B], m.scale, dtype=np.float32)
        dx_o = dx_o + dxi
        assert rel_err(y.detach().float().cpu().numpy(), y_o) < TOL
        assert rel_err(m.downscale_weights[0].grad.float().cpu().numpy(), dA_o[0]) < TOL
        assert rel_err(m.upscale_weights[0].grad.float().cpu().numpy(), dB_o[0]) < TOL
        assert rel_err(m.bias.grad.float().cpu().numpy(), db_o) < TOL
    assert rel_err(x.grad.float().cpu().numpy(), dx_o) < TOL


def test_group_switches_itself_off_when_inputs_differ():
    from tn_gradient.prepare import SoWConfig, prepare_sow

    class Odd(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gate_proj, self.up_proj = torch.nn.Linear(128, 256, bias=False), torch.nn.Linear(128, 256, bias=False)

        def forward(self, x):
            return self.gate_proj(x) + self.up_proj(x * 2.0)       # NOT the same tensor

    torch.manual_seed(0)
    m = prepare_sow(Odd(), SoWConfig(target_modules=["gate_proj", "up_proj"], rank=8, device="cuda", init_method="normal",
                                     decompose="keep")).to("cuda", torch.bfloat16)
    grp = m.gate_proj._group
    x = torch.randn(64, 128, device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        ref = None
        for _ in range(8):
            y = m(x)
            want = (x.float() @ m.gate_proj.acc_downweight.float() + (x.float() @ m.gate_proj.downscale_weights[0].float()) @ m.gate_proj.upscale_weights[0].float()
                    + (2 * x.float()) @ m.up_proj.acc_downweight.float() + ((2 * x.float()) @ m.up_proj.downscale_weights[0].float()) @ m.up_proj.upscale_weights[0].float())
            assert float((y.float() - want).norm() / want.norm()) < 1e-2
    assert not grp.enabled and grp.cache is None


def test_factor_gradients_land_in_the_flat_bucket_without_a_copy():
    """FlatGradSync(direct=factors): the backward kernels write dA / dB straight into the bucket views and autograd adopts
    them as .grad; values equal the ordinary path bit for bit; a second micro-step accumulates."""
    from sow_b200.parallel import FlatGradSync
    blk = _mk_block(256, 512, 8, True)
    ref = _mk_block(256, 512, 8, True)
    ref.load_state_dict(blk.state_dict())
    factors = [p for n, p in blk.named_parameters() if "scale_weights" in n]
    sync = FlatGradSync(factors, overlap=False, direct=factors)
    assert all(p.grad is None for p in factors)
    x = torch.randn(300, 256, device="cuda", dtype=torch.bfloat16)
    blk(x).float().pow(2).mean().backward()
    ref(x).float().pow(2).mean().backward()
    flat = sync.buckets[0]["flat"]
    lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 2
    for (n, p), (_, q) in zip(blk.named_parameters(), ref.named_parameters()):
        if "scale_weights" in n:
            assert lo <= p.grad.data_ptr() < hi and p.grad.data_ptr() == p._sow_grad_view.data_ptr(), n
            assert torch.equal(p.grad, q.grad), n
    blk(x).float().pow(2).mean().backward()                        # micro-step 2: accumulate onto the bucket
    for (n, p), (_, q) in zip(blk.named_parameters(), ref.named_parameters()):
        if "scale_weights" in n:
            assert lo <= p.grad.data_ptr() < hi
            assert float((p.grad.float() - 2 * q.grad.float()).norm() / (2 * q.grad.float().norm() + 1e-20)) < 1e-2, n
    sync.zero_grad()
    assert all(p.grad is None for p in factors) and float(flat.abs().max()) == 0.0


def test_shared_input_groups_stay_grouped_under_torch_compile():
    """Compiled models keep the q/k/v and gate/up groups: dynamo traces the sibling calls into ONE sow_b200::group_fwd node
    per group (zero graph breaks), and the result matches the eager grouped execution."""
    import torch._dynamo
    h, ff, r = 256, 512, 8
    blk = _mk_block(h, ff, r, True)
    x = torch.randn(2, 96, h, device="cuda", dtype=torch.bfloat16)
    xe = x.clone().requires_grad_(True)
    ye = blk(xe)
    ye.float().pow(2).mean().backward()
    ge = {n: p.grad.clone() for n, p in blk.named_parameters() if p.grad is not None}
    gxe = xe.grad.clone()
    blk.zero_grad()
    torch._dynamo.reset()
    ex = torch._dynamo.explain(blk)(x)
    assert ex.graph_break_count == 0, ex.break_reasons
    targets = [str(n.target) for g in ex.graphs for n in g.graph.nodes if n.op == "call_function"]
    assert sum("group_fwd" in t for t in targets) == 2 and sum("linear_fwd" in t for t in targets) == 1, targets
    assert blk.q_proj._group.cache is None and blk.gate_proj._group.cache is None
    xc = x.clone().requires_grad_(True)
    yc = torch.compile(blk, fullgraph=True)(xc)
    yc.float().pow(2).mean().backward()
    # inductor fuses the element-wise glue (silu * up, q + cat(k, v)) in fp32: same function, different bf16 rounding points
    assert float((ye.float() - yc.float()).norm() / ye.float().norm()) < 1e-2
    assert float((gxe.float() - xc.grad.float()).norm() / gxe.float().norm()) < 1e-2
    for n, p in blk.named_parameters():
        if n in ge:
            assert float((ge[n].float() - p.grad.float()).norm() / (ge[n].float().norm() + 1e-20)) < 1e-2, n
