"""GPU parity at the shapes / modes of BASELINE.json's configs 3-5 (the bench line is config 2; config 1 is replayed
in test_optim_loop_gpu.py):

  3. RoBERTa-base-shaped GLUE fine-tune: fp32 modules, biases, decompose='keep', legacy run_glue.py call form
  4. Llama-7B-shaped projections (4096 / 11008, r = 8) under activation checkpointing
  5. TT-QR + TT-Adam at Llama-7B weight shapes (4096x4096, 4096x11008), ranks 8 and 64

Checkers: a plain-torch fp32 restatement of SoWLinear.forward (tn_gradient/layer/sow.py:107-126) run through autograd
on the SAME parameters, the numpy oracle's thin-QR projection, and size-independent identities.  Tolerances are the
north star's: 1e-2 relative for bf16 compute, 1e-5 relative on TT reconstruction errors.
"""
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import sow_oracle as O

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / (b.norm() + 1e-300))


def _torch_sow_linear_group(x, scales, params):
    """sow.py:107-126 in fp32 torch ops (x.W + scale*(x.A).B + bias) for every member of a group; same signature as
    sow_b200.layer.sow_linear_group (per member: W_c, A, B, bias, W_lo -- W = W_c + W_lo for fp32 modules)."""
    xf = x.float()
    outs = []
    for i, scale in enumerate(scales):
        W_c, A, B, bias, W_lo = params[5 * i:5 * i + 5]
        out = scale * ((xf @ A.float()) @ B.float())
        if W_c is not None:
            W = W_c.float() if W_lo is None else W_c.float() + W_lo.float()
            out = out + xf @ W
        if bias is not None:
            out = out + bias.float()
        outs.append(out.to(x.dtype))
    return tuple(outs)


def _run(model, inputs, loss_fn, use_torch, monkeypatch):
    import sow_b200.layer as L
    for p in model.parameters():
        p.grad = None
    if use_torch:
        monkeypatch.setattr(L, "sow_linear_group", _torch_sow_linear_group)
    out = model(**inputs) if isinstance(inputs, dict) else model(inputs)
    logits = out.logits if hasattr(out, "logits") else out
    loss_fn(logits).backward()
    if use_torch:
        monkeypatch.undo()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    return logits.detach().clone(), grads


def test_roberta_glue_shaped_keep_mode_fp32(monkeypatch):
    """Config 3: fp32 RoBERTa-base-shaped encoder (2 of the 12 layers), seq 128 x batch 4, rank 8, mode keep,
    frozen base + trainable classifier, through the legacy call form of scripts/run_glue.py:564-574."""
    from transformers import RobertaConfig, RobertaForSequenceClassification
    from tn_gradient.layer.sow import SoWArgs, SoWLinear
    from tn_gradient.prepare import prepare_sow
    torch.manual_seed(0)
    cfg = RobertaConfig(vocab_size=50265, hidden_size=768, num_hidden_layers=2, num_attention_heads=12,
                        intermediate_size=3072, max_position_embeddings=514, num_labels=2,
                        hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = RobertaForSequenceClassification(cfg)
    targets = ["query", "key", "value", "attention.output.dense", "intermediate.dense", "output.dense"]   # run_glue.py:572
    args = SoWArgs(rank=8, n_iter=1, device="cuda", dtype=torch.float32, scale=0.5, init_method="normal_QR")
    model = prepare_sow(model, targets, decompose="keep", args=args).to("cuda")
    sow = [m for m in model.modules() if isinstance(m, SoWLinear)]
    assert len(sow) == 12 and all(m.bias is not None and m.acc_downweight.dtype == torch.float32 for m in sow)
    assert all(m.acc_downweight.shape == (m.in_features, m.out_features) for m in sow)
    for m in sow:                                             # make the rank-r path matter in the comparison
        with torch.no_grad():
            m.upscale_weights[0].normal_(0, 0.05)
            m.bias.normal_(0, 0.02)
    ids = torch.randint(3, 50264, (4, 128), device="cuda")
    labels = torch.randint(0, 2, (4,), device="cuda")
    loss_fn = lambda logits: nn.functional.cross_entropy(logits.float(), labels)
    y_k, g_k = _run(model, {"input_ids": ids}, loss_fn, False, monkeypatch)
    y_t, g_t = _run(model, {"input_ids": ids}, loss_fn, True, monkeypatch)
    # fp32 modules run the fp32-faithful path (bf16x3 base products, fp32 outputs; the small rank-r factors and their
    # gradients at bf16 input rounding with fp32 accumulation)
    assert _rel(y_k, y_t) < 2e-3
    checked = 0
    for n in g_t:
        if "downscale_weights" in n or "upscale_weights" in n or "classifier" in n:
            assert _rel(g_k[n], g_t[n]) < 1e-2, (n, _rel(g_k[n], g_t[n]))
            checked += 1
    assert checked >= 24
    # biases keep the reference's requires_grad state and receive a gradient when trainable
    n_bias = [n for n in g_t if n.endswith(".bias") and "encoder.layer" in n and "LayerNorm" not in n]
    gmax = max(float(g_t[n].norm()) for n in n_bias)
    for n in n_bias:
        if float(g_t[n].norm()) < 1e-4 * gmax:        # key.bias: softmax is shift-invariant, its true gradient is 0
            assert float(g_k[n].norm()) < 1e-4 * gmax, n
        else:
            assert _rel(g_k[n], g_t[n]) < 2e-2, n


class _Block(nn.Module):
    """Attention-projection + MLP skeleton with the Llama-7B projection shapes (no attention mixing: the SoW path does
    not depend on it)."""

    def __init__(self, h, ff):
        super().__init__()
        self.q_proj, self.k_proj, self.v_proj, self.o_proj = (nn.Linear(h, h, bias=False) for _ in range(4))
        self.gate_proj, self.up_proj = nn.Linear(h, ff, bias=False), nn.Linear(h, ff, bias=False)
        self.down_proj = nn.Linear(ff, h, bias=False)

    def forward(self, x):
        a = self.o_proj(self.q_proj(x) * torch.sigmoid(self.k_proj(x)) + self.v_proj(x))
        x = x + a
        return x + self.down_proj(nn.functional.silu(self.gate_proj(x)) * self.up_proj(x))


def test_llama7b_shaped_block_with_activation_checkpointing(monkeypatch):
    """Config 4: h = 4096, ff = 11008, r = 8, keep mode (fine-tune), bf16, seq 512 x batch 2; forward is re-executed
    inside backward by torch.utils.checkpoint (the C ABI is then entered from the autograd thread)."""
    from torch.utils.checkpoint import checkpoint
    from tn_gradient.prepare import SoWConfig, prepare_sow
    torch.manual_seed(1)
    blk = _Block(4096, 11008)
    for p in blk.parameters():
        nn.init.normal_(p, std=0.02)
    targets = ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"]
    blk = prepare_sow(blk, SoWConfig(target_modules=targets, rank=8, scale=0.125, device="cuda",
                                     init_method="normal_QR", decompose="keep")).to("cuda", torch.bfloat16)
    with torch.no_grad():
        for n, p in blk.named_parameters():
            if "upscale_weights" in n:
                p.normal_(0, 0.05)
    x = (torch.randn(2, 512, 4096, device="cuda") * 0.5).bfloat16().requires_grad_(True)
    loss_fn = lambda y: y.float().pow(2).mean()

    def run(fn):
        x.grad = None
        for p in blk.parameters():
            p.grad = None
        loss_fn(fn(x)).backward()
        return x.grad.clone(), {n: p.grad.clone() for n, p in blk.named_parameters() if p.grad is not None}

    gx_plain, g_plain = run(blk)
    gx_ckpt, g_ckpt = run(lambda t: checkpoint(blk, t, use_reentrant=False))
    assert torch.equal(gx_plain, gx_ckpt)
    assert len(g_plain) == 14                                                                    # only factors train
    for n in g_plain:        # split-K partials are summed in a fixed order: bit-reproducible
        assert torch.equal(g_ckpt[n], g_plain[n]), n
    import sow_b200.layer as L
    monkeypatch.setattr(L, "sow_linear_group", _torch_sow_linear_group)
    gx_t, g_t = run(blk)
    monkeypatch.undo()
    assert _rel(gx_plain, gx_t) < 1e-2
    for n in g_t:
        assert _rel(g_plain[n], g_t[n]) < 1e-2, (n, _rel(g_plain[n], g_t[n]))


def _interleave_np(mat, mm, nn_):
    M, N = mat.shape
    pad = np.zeros((mm * mm, nn_ * nn_), dtype=mat.dtype)
    pad[:M, :N] = mat
    return pad.reshape(mm, mm, nn_, nn_).transpose(0, 2, 1, 3).reshape(mm * nn_, mm * nn_)      # (i1,o1 | i2,o2)


@pytest.mark.parametrize("M,N,r", [(4096, 4096, 8), (4096, 4096, 64), (4096, 11008, 16)])
def test_tt_decomposition_at_llama7b_shapes_vs_projection_oracle(M, N, r):
    """Config 5: order-2 TT of a Llama-7B-shaped matrix.  The reference's complete QR + truncation (tt.py:127-136) is the
    orthogonal projection onto the span of the first r columns of the unfolding (checked against the reference at small
    sizes in test_from_matrix_to_matrix_vs_reference); at this size the oracle evaluates that projection directly."""
    from math import ceil
    from tn_gradient.tt import TensorTrain
    rng = np.random.default_rng(5)
    low = rng.standard_normal((M, 24)).astype(np.float32) @ rng.standard_normal((24, N)).astype(np.float32)
    mat_np = (low / 24 ** 0.5 + 0.3 * rng.standard_normal((M, N)).astype(np.float32)).astype(np.float32)
    mat = torch.from_numpy(mat_np).cuda()
    tt = TensorTrain.from_matrix(mat, [1, r, 1])
    back = tt.to_matrix(mat.shape)
    err = float((back - mat).norm() / mat.norm())
    mm, nn_ = ceil(M ** 0.5), ceil(N ** 0.5)
    L = _interleave_np(mat_np.astype(np.float64), mm, nn_)
    Q = O.thin_q(L, r)
    R = Q.T @ L
    ref_err = float(np.sqrt(max(np.sum(L * L) - np.sum(R * R), 0.0)) / np.linalg.norm(mat_np.astype(np.float64)))
    # the padded region of the reconstruction is cut away by to_matrix, so err <= ref_err; equal when nothing is padded
    if mm * mm == M and nn_ * nn_ == N:
        assert abs(err - ref_err) <= 1e-5 * ref_err, (err, ref_err)
    else:
        assert err <= ref_err * (1 + 1e-5) and err > 0.5 * ref_err, (err, ref_err)
    G1 = tt.cores[0].reshape(mm * nn_, r)
    assert float((G1.T @ G1 - torch.eye(r, device="cuda")).abs().max()) < 1e-5


@pytest.mark.parametrize("M,N,r,dtype", [(4096, 4096, 64, torch.bfloat16), (4096, 11008, 8, torch.float32),
                                         (4096, 4096, 32, torch.float32), (1024, 4096, 48, torch.float32)])
def test_ttadam_fused_step_at_llama7b_shapes(M, N, r, dtype):
    """Config 5: two TTAdam steps (ttadam.py:61-115).  Step 1 starts from zero moments, so the parameter update is the
    dense Adam update; step 2 must use the reconstruction of the compressed moments, and the new compressed moments
    must be the projection of the dense ones (checked through the identity |m2|^2 = |Q^T m2|^2 + |m2 - QQ^T m2|^2)."""
    from tn_gradient.optimizer.ttadam import TTAdam
    torch.manual_seed(2)
    # bf16 parameters: the step must be large against ulp(p) for the update itself to be observable
    lr, b1, b2, eps = (5e-2 if dtype == torch.bfloat16 else 1e-3), 0.9, 0.999, 1e-8
    p = nn.Parameter((torch.randn(M, N, device="cuda") * 0.02).to(dtype))
    opt = TTAdam([{"params": [p], "ranks": [1, r, 1]}], lr=lr, betas=(b1, b2), eps=eps)
    tol = 2e-2 if dtype == torch.bfloat16 else 2e-5
    from math import ceil
    padded = ceil(M ** 0.5) ** 2 != M or ceil(N ** 0.5) ** 2 != N

    def dense_update(p0, g, m_prev, v_prev, step):
        g = g.float()
        m = b1 * m_prev + (1 - b1) * g
        v = b2 * v_prev.clamp_min(0) + (1 - b2) * g * g
        step_size = lr * (1 - b2 ** step) ** 0.5 / (1 - b1 ** step)
        return p0.float() - step_size * m / (v.sqrt() + eps), m, v

    zeros = torch.zeros(M, N, device="cuda")
    g1 = (torch.randn(M, N, device="cuda") * 0.01).to(dtype)
    p0 = p.detach().clone()
    p.grad = g1.clone()
    opt.step()
    want1, m1, v1 = dense_update(p0, g1, zeros, zeros, 1)
    assert _rel(p.detach().float() - p0.float(), want1 - p0.float()) < tol
    st = opt.state[p]
    m1_tt, v1_tt = st["exp_avg"].to_matrix((M, N)), st["exp_avg_sq"].to_matrix((M, N))
    # compressed moments = orthogonal projection of the dense ones: <m1 - m1_tt, m1_tt> = 0 up to rounding (exact when
    # nothing is padded; padding only removes rows/columns of the residual)
    for dense, comp in ((m1, m1_tt), (v1, v1_tt)):
        assert comp.shape == dense.shape
        cos = float(((dense - comp) * comp).sum() / (dense.norm() * comp.norm() + 1e-30))
        assert abs(cos) < (3e-2 if padded else 1e-4), cos
        assert float(comp.norm()) <= float(dense.norm()) * (1 + 1e-5)
    g2 = (torch.randn(M, N, device="cuda") * 0.01).to(dtype)
    p1 = p.detach().clone()
    p.grad = g2.clone()
    opt.step()
    want2, _, _ = dense_update(p1, g2, m1_tt, v1_tt, 2)
    assert _rel(p.detach().float() - p1.float(), want2 - p1.float()) < tol
    assert st is opt.state[p] and opt.state[p]["step"] == 2


def test_decompose_qr_surgery_preserves_the_layer_function():
    """prepare_sow(decompose='qr') (prepare.py:122-147): W^T = Q R is split into a frozen major part Q[:, :-r] R[:-r, :] and the
    trainable minor factors Q[:, -r:], R[-r:, :]; with scale = 1 the SoW layer reproduces the dense layer it replaced."""
    from tn_gradient.layer.sow import SoWLinear
    from tn_gradient.prepare import SoWConfig, prepare_sow
    torch.manual_seed(4)
    dense = nn.Sequential(nn.Linear(512, 1376, bias=True), nn.Linear(1376, 512, bias=False)).to("cuda")
    x = torch.randn(256, 512, device="cuda")
    with torch.no_grad():
        want = [dense[0](x), dense[1](torch.tanh(dense[0](x)))]
    model = prepare_sow(dense, SoWConfig(target_modules=["0", "1"], rank=8, scale=1.0, device="cuda", decompose="qr"))
    assert all(isinstance(m, SoWLinear) for m in model)
    for m in model:
        assert m.acc_downweight.shape == (m.in_features, m.out_features) and not m.acc_downweight.requires_grad
        assert m.downscale_weights[0].shape == (m.in_features, 8) and m.upscale_weights[0].shape == (8, m.out_features)
        A = m.downscale_weights[0].float()
        assert float((A.T @ A - torch.eye(8, device="cuda")).abs().max()) < 1e-4        # columns of Q
    with torch.no_grad():
        got = [model[0](x), model[1](torch.tanh(model[0](x)))]
    for g, w in zip(got, want):
        assert _rel(g, w) < 1e-2                                                        # bf16 compute policy on fp32 modules
    # the minor factors train, the major part does not
    model[1](torch.tanh(model[0](x))).pow(2).mean().backward()
    assert model[0].downscale_weights[0].grad is not None and model[0].acc_downweight.grad is None
