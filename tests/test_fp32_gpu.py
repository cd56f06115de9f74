"""GPU parity of the fp32-faithful path (fp32 modules: the reference's GLUE fine-tunes never set a dtype, run_glue.py:
386-388,508-514): bf16x3 tensor-core base products with fp32 outputs, and the exact-fp32 grouped merge that keeps the
pretrained weights at full precision (sow.py:131-153 accumulates in the parameter dtype)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import sow_oracle as O

pytestmark = pytest.mark.gpu


def _layer(fin, fout, r, scale, bias, seed=0):
    from tn_gradient.layer.sow import SoWLinear
    torch.manual_seed(seed)
    layer = SoWLinear(fin, fout, bias=bias, rank=r, scale=scale, init_method="normal", dtype=torch.float32, device="cuda")
    layer.virtual_rank = min(fin, fout)
    layer.acc_downweight = torch.nn.Parameter(torch.randn(fin, fout, device="cuda") * 0.02, requires_grad=False)
    with torch.no_grad():
        layer.downscale_weights[0].normal_(0, 0.05)
        layer.upscale_weights[0].normal_(0, 0.05)
        if bias:
            layer.bias.normal_(0, 0.1)
    return layer


@pytest.mark.parametrize("T,fin,fout,r,scale,bias", [(1000, 768, 768, 8, 0.125, True), (512, 768, 3072, 8, 0.125, True),
                                                     (300, 264, 136, 50, 1.0, False), (2048, 1024, 2736, 50, 1.0, False)])
def test_fp32_module_forward_backward_is_fp32_faithful(T, fin, fout, r, scale, bias):
    layer = _layer(fin, fout, r, scale, bias)
    x = torch.randn(T, fin, device="cuda", requires_grad=True)
    dy = torch.randn(T, fout, device="cuda")
    y = layer(x)
    assert y.dtype == torch.float32
    y.backward(dy)
    torch.cuda.synchronize()
    W, A, B = (t.detach().cpu().numpy() for t in (layer.acc_downweight, layer.downscale_weights[0], layer.upscale_weights[0]))
    b = layer.bias.detach().cpu().numpy() if bias else None
    xn, dyn = x.detach().cpu().numpy(), dy.cpu().numpy()
    y_o = O.sow_linear_forward(xn, W, [A], [B], b, scale)                                 # fp64 truth
    dx_o, dA_o, dB_o, db_o = O.sow_linear_backward(dyn, xn, W, [A], [B], scale)
    # the base term (x.W, dY.W^T) is fp32-faithful; the rank-r term enters with bf16-rounded factors, which bounds the
    # total error by its relative weight in the output
    y_lr = scale * (xn.astype(np.float64) @ A) @ B
    dx_lr = scale * (dyn.astype(np.float64) @ B.T) @ A.T
    w_y = float(np.linalg.norm(y_lr) / np.linalg.norm(y_o))          # weight of the rank-r term in y / dX
    w_dx = float(np.linalg.norm(dx_lr) / np.linalg.norm(dx_o))
    assert rel_err(y.detach().cpu().numpy(), y_o) < 1e-4 + 4e-3 * w_y
    assert rel_err(x.grad.cpu().numpy(), dx_o) < 1e-4 + 4e-3 * w_dx
    assert rel_err(layer.downscale_weights[0].grad.cpu().numpy(), dA_o[0]) < 1e-2         # bf16 inputs, fp32 accumulate
    assert rel_err(layer.upscale_weights[0].grad.cpu().numpy(), dB_o[0]) < 1e-2
    if bias:
        assert rel_err(layer.bias.grad.cpu().numpy(), db_o) < 1e-2
    # and against plain bf16 compute the improvement is real: the same layer in bf16 is two orders of magnitude off
    lb = _layer(fin, fout, r, scale, bias).to(torch.bfloat16)
    lb.acc_downweight = torch.nn.Parameter(layer.acc_downweight.detach().bfloat16(), requires_grad=False)
    with torch.no_grad():
        lb.downscale_weights[0].copy_(layer.downscale_weights[0])
        lb.upscale_weights[0].copy_(layer.upscale_weights[0])
        if bias:
            lb.bias.copy_(layer.bias)
        e_bf16 = rel_err(lb(x.detach().bfloat16()).float().cpu().numpy(), y_o)
    if w_y < 0.2:
        assert rel_err(y.detach().cpu().numpy(), y_o) < 0.3 * e_bf16


@pytest.mark.parametrize("fin,fout,r,scale,has_prev", [(768, 768, 8, 0.125, True), (768, 3072, 8, 0.125, True),
                                                       (3072, 768, 8, 1.0, False), (100, 136, 7, 0.5, True),
                                                       (264, 130, 50, 1.0, True), (128, 256, 100, 1.0, True)])
def test_fp32_merge_is_exact_to_fp32_rounding_and_in_place(fin, fout, r, scale, has_prev):
    layer = _layer(fin, fout, r, scale, False)
    if not has_prev:
        layer.acc_downweight = torch.nn.Parameter(torch.empty(0, device="cuda"), requires_grad=False)
    W = layer.acc_downweight.detach().cpu().numpy() if has_prev else None
    A, B = layer.downscale_weights[0].detach().cpu().numpy(), layer.upscale_weights[0].detach().cpu().numpy()
    ptr = layer.acc_downweight.data_ptr()
    x = torch.randn(64, fin, device="cuda")
    fwd_ok = fin % 8 == 0 and fout % 8 == 0               # the GEMMs need 16-byte row pitches; the fp32 merge does not
    if fwd_ok and has_prev:
        with torch.no_grad():
            layer(x)                                      # builds the split shadow of W: must be dropped by the merge
    layer.accumulate()
    torch.cuda.synchronize()
    W_o, _, _ = O.sow_merge(W, [A], [B], scale)                                            # fp64
    got = layer.acc_downweight.detach().cpu().numpy()
    assert layer.acc_downweight.dtype == torch.float32 and layer.acc_upweight.numel() == 0
    assert rel_err(got, W_o) < 1e-6
    if has_prev:
        assert layer.acc_downweight.data_ptr() == ptr     # in place, no bf16 round trip of the pretrained weights
    if fwd_ok:
        with torch.no_grad():
            y = layer(x)                                  # forward after the merge reads the NEW W (B is zero now)
        assert rel_err(y.cpu().numpy(), x.cpu().numpy().astype(np.float64) @ W_o) < 1e-4
