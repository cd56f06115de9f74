"""Debug aid: per-tile pipeline timeline of CTA 0 of the grouped merge (clock64 stamps)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sow_b200 import ops, _lib
sys.argv = [sys.argv[0]] + sys.argv[1:]
from tools.bench_merge import layer_shapes
name = sys.argv[1] if len(sys.argv) > 1 else "llama_350m"
r = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda", 0)
items = []
for fin, fout in layer_shapes(name):
    W = (torch.randn(fin, fout, device=dev) * 0.02).bfloat16()
    items.append((W, W, (torch.randn(fin, r, device=dev) * 0.05).bfloat16(), (torch.randn(r, fout, device=dev) * 0.05).bfloat16(), 1.0))
for _ in range(3):
    ops.merge_grouped(items)
torch.cuda.synchronize()
ts = torch.zeros(256 * 8 + 2 * 160, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.sow_merge_debug_timeline(ctypes.c_void_p(ts.data_ptr()))
ops.merge_grouped(items)
torch.cuda.synchronize()
lib.sow_merge_debug_timeline(None)
cta = ts[2048:].view(160, 2).cpu()
t = ts[:2048].view(256, 8).cpu()
n = int((t[:, 0] != 0).sum())
t = t[:n] - t[0, 0]
names = ["prod_reach", "slot_free", "epi_reach", "tfull", "W_landed", "epi_done", "store_issue", "store_read_done"]
print("tile " + " ".join(f"{x:>12s}" for x in names))
for i in list(range(0, min(n, 24))) + list(range(max(24, n - 6), n)):
    print(f"{i:4d} " + " ".join(f"{int(v):12d}" for v in t[i]))
import numpy as np
a = t.numpy().astype(np.float64)
print("tiles", n, "total cycles", a[-1, 7])
mid = a[8:-4]
print("avg cycles/tile (steady):", (mid[-1, 7] - mid[0, 7]) / (len(mid) - 1))
print("avg load latency (slot_free -> W_landed):", np.mean(mid[:, 4] - mid[:, 1]))
print("avg epilogue (W_landed -> epi_done):", np.mean(mid[:, 5] - mid[:, 4]))
print("avg epi wait for tfull (epi_reach -> tfull):", np.mean(mid[:, 3] - mid[:, 2]))
print("avg epi wait for W (tfull -> W_landed):", np.mean(mid[:, 4] - mid[:, 3]))
print("avg store (store_issue -> read_done):", np.mean(mid[:, 7] - mid[:, 6]))
print("avg producer wait for slot:", np.mean(mid[:, 1] - mid[:, 0]))

live = cta[cta[:, 0] > 0]
t0 = int(live[:, 0].min())
ends = sorted((live[:, 1] - t0).tolist())
starts = sorted((live[:, 0] - t0).tolist())
print(f"per-CTA wall time: {len(live)} CTAs, start spread {starts[-1]} ns, end min/median/max = {ends[0]}/{ends[len(ends)//2]}/{ends[-1]} ns")
