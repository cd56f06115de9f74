"""Debug aid: are the slow CTAs of the grouped merge the same ones launch after launch?  Per-CTA wall time (globaltimer) of
several launches, their rank correlation, and the SM each CTA ran on (from the kernel's debug stamps)."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sow_b200 import ops, _lib
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
lib = _lib.load()
durs = []
for rep in range(6):
    ts = torch.zeros(256 * 8 + 2 * 160, dtype=torch.int64, device=dev)
    lib.sow_merge_debug_timeline(ctypes.c_void_p(ts.data_ptr()))
    if rep % 2:                                       # something else in front of every other launch
        x = torch.randn(4096, 4096, device=dev) @ torch.randn(4096, 4096, device=dev)
    ops.merge_grouped(items)
    torch.cuda.synchronize()
    lib.sow_merge_debug_timeline(None)
    cta = ts[2048:].view(160, 2).cpu().numpy()
    live = cta[:, 0] > 0
    d = (cta[live, 1] - cta[live, 0]).astype(np.float64)
    durs.append(d)
    print(f"launch {rep}: {live.sum()} CTAs, duration min/median/max = {d.min():.0f}/{np.median(d):.0f}/{d.max():.0f} ns, span {cta[live,1].max()-cta[live,0].min()} ns")
D = np.stack(durs)
C = np.corrcoef(D)
print("correlation of per-CTA durations between launches:")
print(np.round(C, 2))
mean = D.mean(0)
print("mean duration by CTA, sorted: slowest 10 CTAs", np.argsort(-mean)[:10].tolist(), "fastest 10", np.argsort(mean)[:10].tolist())
print("relative spread of the mean: min %.3f max %.3f of the average" % (mean.min() / mean.mean(), mean.max() / mean.mean()))
