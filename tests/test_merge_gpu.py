"""GPU parity: SoWLinear.accumulate / tn_gradient.prepare.accumulate (grouped merge kernel + thin-QR re-init)
against the oracle and the reference's golden outputs."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import sow_oracle as O

pytestmark = pytest.mark.gpu


def _mk(fin, fout, r, n_iter, scale, dtype, init="normal"):
    from tn_gradient.layer.sow import SoWLinear
    return SoWLinear(fin, fout, bias=False, rank=r, n_iter=n_iter, scale=scale, init_method=init, dtype=dtype, device="cuda")


@pytest.mark.parametrize("case", ["dense_f32", "dense_bf16", "dense_niter2_f32", "factored_f32", "factored_bf16"])
def test_merge_matches_reference_golden(golden_merge, case):
    g = golden_merge
    fin, fout, r, n_iter = [int(v) for v in g[f"merge/{case}/meta"]]
    scale = float(g[f"merge/{case}/scale"])
    dtype = torch.bfloat16 if "bf16" in case else torch.float32
    layer = _mk(fin, fout, r, n_iter, scale, dtype)
    if not case.startswith("factored"):
        layer.virtual_rank = min(fin, fout)
    ids = [id(p) for p in list(layer.downscale_weights) + list(layer.upscale_weights)]
    for step in range(3):
        p = f"merge/{case}/step{step}/"
        with torch.no_grad():
            for i in range(n_iter):
                layer.downscale_weights[i].copy_(torch.from_numpy(g[p + f"A{i}"]))
                layer.upscale_weights[i].copy_(torch.from_numpy(g[p + f"B{i}"]))
        if case.startswith("factored") and step > 0:
            # keep the accumulated factors on the reference's trajectory (gauge differs otherwise)
            layer.acc_downweight = torch.nn.Parameter(torch.from_numpy(g[p + "W_prev"]).cuda(), requires_grad=False)
            layer.acc_upweight = torch.nn.Parameter(torch.from_numpy(g[p + "Wup_prev"]).cuda(), requires_grad=False)
        layer.accumulate()
        torch.cuda.synchronize()
        assert layer.virtual_rank == int(g[p + "virtual_rank"])
        Wd = layer.acc_downweight.float().cpu().numpy()
        if g[p + "Wup"].size == 0:
            assert layer.acc_upweight.numel() == 0
            assert rel_err(Wd, g[p + "W"]) < 1e-2           # bf16 compute policy, fp32 accumulate
            assert float(np.abs(Wd).max()) > 0                # the merged update was kept (not written to a dropped temporary)
        else:
            Wu = layer.acc_upweight.float().cpu().numpy()
            assert Wd.shape == g[p + "W"].shape and Wu.shape == g[p + "Wup"].shape
            assert rel_err(Wd @ Wu, g[p + "W"].astype(np.float64) @ g[p + "Wup"].astype(np.float64)) < (3e-2 if "bf16" in case else 1e-4)
        # factor re-init: B zero, A ~ N(0, .02) here (init_method "normal"), Parameter identity preserved
        assert [id(q) for q in list(layer.downscale_weights) + list(layer.upscale_weights)] == ids
        for B in layer.upscale_weights:
            assert float(B.abs().max()) == 0.0
        for A in layer.downscale_weights:
            assert 0.01 < float(A.float().std()) < 0.04
        assert not layer.acc_downweight.requires_grad


@pytest.mark.parametrize("fin,fout,r,scale,has_prev", [
    (1024, 1024, 50, 1.0, True), (1024, 2736, 50, 1.0, True), (2736, 1024, 50, 0.5, False),
    (768, 3072, 8, 0.125, True), (4096, 11008, 8, 0.125, True), (264, 136, 50, 1.0, True), (128, 128, 200, 1.0, True),
    (100, 136, 7, 1.0, True), (204, 128, 33, 0.5, False), (1000, 264, 64, 1.0, True), (384, 2736, 16, 2.0, True)])
def test_merge_vs_oracle_at_baseline_shapes(fin, fout, r, scale, has_prev):
    rng = np.random.default_rng(7)
    A = O.bf16_round(rng.standard_normal((fin, r), dtype=np.float32) * 0.05)
    B = O.bf16_round(rng.standard_normal((r, fout), dtype=np.float32) * 0.05)
    W = O.bf16_round(rng.standard_normal((fin, fout), dtype=np.float32) * 0.02) if has_prev else None
    layer = _mk(fin, fout, r, 1, scale, torch.bfloat16)
    layer.virtual_rank = min(fin, fout)
    with torch.no_grad():
        layer.downscale_weights[0].copy_(torch.from_numpy(A))
        layer.upscale_weights[0].copy_(torch.from_numpy(B))
    if has_prev:
        layer.acc_downweight = torch.nn.Parameter(torch.from_numpy(W).to("cuda", torch.bfloat16), requires_grad=False)
        ptr = layer.acc_downweight.data_ptr()
    layer.accumulate()
    torch.cuda.synchronize()
    W_o, _, _ = O.sow_merge(W, [A], [B], scale, dtype=np.float32)
    got = layer.acc_downweight.float().cpu().numpy()
    assert rel_err(got, W_o) < 4e-3                      # single bf16 rounding of an fp32 accumulation
    if has_prev:
        assert layer.acc_downweight.data_ptr() == ptr    # merged in place: pointer-stable for graphs / TMA maps


def test_grouped_accumulate_over_a_model_and_forward_continuity():
    """prepare.accumulate(model): one grouped launch; the layer function is unchanged by a merge (B_new = 0):
    forward(after) == forward(before) up to bf16 rounding (size-independent property)."""
    import torch.nn as nn
    from tn_gradient.prepare import SoWConfig, accumulate, prepare_sow
    torch.manual_seed(3)
    model = nn.Sequential(nn.Linear(512, 1376, bias=False), nn.Linear(1376, 512, bias=True), nn.Linear(512, 512, bias=False))
    prepare_sow(model, SoWConfig(target_modules=["0", "1", "2"], rank=50, scale=1.0, device="cuda",
                                 init_method="normal_QR", decompose="keep"))
    model = model.to("cuda", torch.bfloat16)
    for m in model:
        with torch.no_grad():
            m.upscale_weights[0].normal_(0, 0.02)
    x = torch.randn(512, 512, device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        before = [m(x if m.in_features == 512 else torch.randn(512, m.in_features, device="cuda", dtype=torch.bfloat16).mul_(0) + 1) for m in model]
        accumulate(model)
        after = [m(x if m.in_features == 512 else torch.ones(512, m.in_features, device="cuda", dtype=torch.bfloat16)) for m in model]
    for b, a in zip(before, after):
        assert float((a.float() - b.float()).norm() / b.float().norm()) < 1e-2
    for m in model:
        A = m.downscale_weights[0].float()
        assert float((A.T @ A - torch.eye(50, device="cuda")).abs().max()) < 2e-2     # orthonormal up to bf16 storage
        assert float(m.upscale_weights[0].abs().max()) == 0.0


def test_reinit_normal_qr_spans_reference_subspace(golden_merge):
    """thin-QR kernel on the reference's own Gaussian draw: same subspace as the reference's A_new, and equal to the
    oracle's sign-normalised Q."""
    from sow_b200 import ops
    g = golden_merge
    G = torch.from_numpy(g["reinit/gauss"]).cuda()
    r = g["reinit/A_new"].shape[1]
    Q = ops.thin_qr(G, r).cpu().numpy()
    assert rel_err(Q, O.thin_q(g["reinit/gauss"], r, positive_diag=True)) < 1e-5
    A_ref = g["reinit/A_new"].astype(np.float64)
    assert rel_err(Q @ (Q.T @ A_ref), A_ref) < 1e-5           # span(Q) contains the reference's columns
    assert np.abs(np.abs(np.diag(Q.T @ A_ref)) - 1).max() < 1e-5   # column-wise equal up to sign


def test_qr_weight_kernel_path_vs_reference(golden_merge):
    from tn_gradient.utils import qr_weight
    g = golden_merge
    for case in ["tall", "wide", "square"]:
        w = torch.from_numpy(g[f"qr/{case}/w"]).cuda()
        r = int(g[f"qr/{case}/rank"])
        Q, R = qr_weight(w, r)
        Qn, Rn = Q.cpu().numpy().astype(np.float64), R.cpu().numpy().astype(np.float64)
        assert rel_err(Qn @ Rn, g[f"qr/{case}/Q"].astype(np.float64) @ g[f"qr/{case}/R"].astype(np.float64)) < 1e-5
        s = np.sign(np.diag(Qn.T @ g[f"qr/{case}/Q"]))
        assert rel_err(Qn * s, g[f"qr/{case}/Q"]) < 1e-5
        assert rel_err(Rn * s[:, None], g[f"qr/{case}/R"]) < 1e-5


@pytest.mark.parametrize("m,r,noise", [(4096, 16, 1e-4), (4096, 64, 3e-5), (1024, 8, 1e-5)])
def test_thin_qr_on_ill_conditioned_positive_columns(m, r, noise):
    """The leading columns of a TT-Adam second-moment unfolding are all positive and nearly rank one (cond ~ 1e4-1e6): the
    single-pass Cholesky-QR loses orthogonality as cond^2, so a second Cholesky-QR pass (CholeskyQR2) runs when a pivot ratio
    trips.  Q must stay orthonormal to fp32 level and span the same space as LAPACK's Householder Q."""
    from sow_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    u = torch.rand(m, 1, device="cuda", generator=g) + 0.5
    v = torch.rand(1, r, device="cuda", generator=g) + 0.5
    X = (u @ v) * (1.0 + noise * torch.rand(m, r, device="cuda", generator=g))      # positive, nearly rank one
    Q = ops.thin_qr(X, r)
    torch.cuda.synchronize()
    Qd = Q.double()
    ortho = float((Qd.T @ Qd - torch.eye(r, device="cuda", dtype=torch.float64)).abs().max())
    assert ortho < 2e-6, ortho
    Q_ref, _ = torch.linalg.qr(X.double())
    resid = float((Q_ref - Qd @ (Qd.T @ Q_ref)).norm() / Q_ref.norm())                # span(Q_ref) inside span(Q)
    assert resid < 2e-3, resid
    # the well-conditioned case still takes the fast path and agrees with LAPACK to rounding
    G = torch.randn(m, r, device="cuda", generator=g)
    Qg = ops.thin_qr(G, r).double()
    Qr, Rr = torch.linalg.qr(G.double())
    Qr = Qr * torch.sign(torch.diagonal(Rr))
    assert float((Qg - Qr).abs().max()) < 1e-5
