"""GPU: the fall-back kernels of the tensor-train path (tiled projections / reconstructions, the shared-memory TT-Adam kernel,
scalar fp64 Gram / solve) stay equivalent to the default ones.  The switches (SOWB_TT_REG, SOWB_QR_MMA) are read once per
process, so each setting runs in its own interpreter and the results are compared here."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

SCRIPT = r"""
import sys, numpy as np, torch
sys.path.insert(0, %(root)r)
from tn_gradient.tt import TensorTrain
from tn_gradient.optimizer.ttadam import TTAdam
from sow_b200 import ops
out = {}
torch.manual_seed(7)
mat = torch.randn(1024, 1536, device="cuda")
for order, r in ((2, 8), (2, 24), (2, 48), (3, 8)):
    ranks = [1] + [r] * (order - 1) + [1]
    tt = TensorTrain.from_matrix(mat, ranks)
    out[f"recon_o{order}_r{r}"] = tt.to_matrix(mat.shape).cpu().numpy()
X = torch.randn(3, 2048, 64, device="cuda")
for r in (8, 24, 50, 64):
    Q = ops.thin_qr(X, r)
    out[f"qtq_r{r}"] = (Q.transpose(1, 2) @ Q).cpu().numpy()
    out[f"proj_r{r}"] = (Q @ (Q.transpose(1, 2) @ X[:, :, :r])).cpu().numpy()       # sign / basis independent
for r in (8, 16, 40):
    torch.manual_seed(9)
    p = torch.nn.Parameter(torch.randn(1024, 1024, device="cuda"))
    opt = TTAdam([{"params": [p], "ranks": [1, r, 1]}], lr=1e-2)
    for step in range(3):
        p.grad = torch.randn(1024, 1024, device="cuda", generator=torch.Generator("cuda").manual_seed(step))
        opt.step()
    out[f"p_r{r}"] = p.detach().cpu().numpy()
    out[f"m_r{r}"] = opt.state[p]["exp_avg"].to_matrix(p.shape).cpu().numpy()
np.savez(sys.argv[1], **out)
"""


def run(env_extra, path):
    env = dict(os.environ)
    env.update(env_extra)
    subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}, path], check=True, env=env, timeout=600)
    return dict(np.load(path))


def test_fallback_kernels_match_the_default_ones():
    with tempfile.TemporaryDirectory() as d:
        ref = run({}, os.path.join(d, "default.npz"))
        alt = run({"SOWB_TT_REG": "0", "SOWB_QR_MMA": "0"}, os.path.join(d, "fallback.npz"))
    assert sorted(ref) == sorted(alt)
    for key in ref:
        a, b = ref[key].astype(np.float64), alt[key].astype(np.float64)
        err = np.linalg.norm(a - b) / max(np.linalg.norm(a), 1e-30)
        # same algorithm, different summation orders (and MUFU vs IEEE sqrt / divide inside the Adam update)
        assert err < 2e-5, (key, err)
