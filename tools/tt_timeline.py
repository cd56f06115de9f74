"""Debug aid: per-tile phase timeline of CTA (0,0) of the tensor-core TT-Adam kernel."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sow_b200 import _lib
from tn_gradient.optimizer.ttadam import TTAdam
M, N, r = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (4096, 4096, 64))]
dev = torch.device("cuda", 0)
p = torch.nn.Parameter((torch.randn(M, N, device=dev) * 0.02).bfloat16())
p.grad = (torch.randn(M, N, device=dev) * 0.01).bfloat16()
opt = TTAdam([{"params": [p], "ranks": [1, r, 1]}], lr=1e-3)
for _ in range(3):
    opt.step()
torch.cuda.synchronize()
ts = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.tt_adam2_debug_timeline(ctypes.c_void_p(ts.data_ptr()))
opt.step()
torch.cuda.synchronize()
lib.tt_adam2_debug_timeline(None)
t = ts.view(64, 8).cpu()
n = int((t[:, 0] != 0).sum())
t = t[:n, :6] - t[0, 0]
print("tile   start  S_ready  epiA_done  Dm_done  epiB_done  Dv_done")
for i in range(n):
    print(i, [int(v) for v in t[i]])
d = t[:, 1:] - t[:, :-1]
print("avg phase cycles: phase1(load+MMA S) %.0f | epilogue A %.0f | MMA D_m %.0f | epilogue B %.0f | MMA D_v %.0f" % tuple(d.float().mean(0).tolist()))
