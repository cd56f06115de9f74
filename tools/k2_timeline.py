"""Debug: clock64 timeline of CTA 0 of the K2 kernel (one launch), per T-chunk.  python tools/k2_timeline.py [T out]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sow_b200 import _lib, ops  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
out = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
nm = int(sys.argv[3]) if len(sys.argv) > 3 else 1
fin, r = 1024, 50
dev = torch.device("cuda", 0)
x = torch.randn(T, fin, device=dev).bfloat16()
mem = [((torch.randn(fin, out, device=dev) * 0.02).bfloat16(), (torch.randn(fin, r, device=dev) * 0.05).bfloat16(),
        (torch.randn(r, out, device=dev) * 0.05).bfloat16(), torch.randn(T, out, device=dev).bfloat16()) for _ in range(nm)]
ys, A_cat, t_cat = ops.group_fwd(x, [(W, A, B, None, 1.0) for W, A, B, dy in mem])
bw = [(W, B, dy, 1.0, True, True, False) for W, A, B, dy in mem]
for _ in range(2):
    ops.group_bwd(x, A_cat, t_cat, bw, False)
ts = torch.zeros(64 * 16 + 512 * 3, dtype=torch.int64, device=dev)
lib = ctypes.CDLL(_lib.LIB_PATH)
lib.sow_k2_debug_timeline(ctypes.c_void_p(ts.data_ptr()))
ops.group_bwd(x, A_cat, t_cat, bw, False)
torch.cuda.synchronize()
lib.sow_k2_debug_timeline(None)
full = ts.cpu()
t = full[:1024].view(64, 16)
t0 = int(t[0, 0])
names = ["prod:start", "-", "prod:issued", "mma:start", "-", "mma:ready", "mma:issued",
         "epi:start", "epi:acc in regs", "epi:sent", "epi:summed", "epi:done"]
for c in range(16):
    if int(t[c, 0]) == 0:
        break
    print(f"chunk {c}: " + "  ".join(f"{n}={int(t[c, k]) - t0}" for k, n in enumerate(names) if int(t[c, k]) and n != "-"))
print("final dB: " + "  ".join(f"{n}={int(t[63, k]) - t0}" for k, n in enumerate(["wait", "db_full", "written"]) if int(t[63, k])))

cta = full[1024:].view(512, 3)
live = cta[cta[:, 0] > 0]
if len(live):
    t_begin = int(live[:, 0].min())
    starts = (live[:, 0] - t_begin).tolist()
    ends = (live[:, 1] - t_begin).tolist()
    print(f"CTAs: {len(live)}  distinct SMs: {len(set(live[:, 2].tolist()))}  start ns min/median/max = {min(starts)}/{sorted(starts)[len(starts)//2]}/{max(starts)}  "
          f"end ns min/median/max = {min(ends)}/{sorted(ends)[len(ends)//2]}/{max(ends)}")
    late = [(i, s) for i, s in enumerate(starts) if s > 5000]
    print("CTAs starting > 5 us after the first:", len(late), late[:12])
