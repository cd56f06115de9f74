import os, sys, torch
sys.path.insert(0, "/root/repo")
from sow_b200 import ops
dev = torch.device("cuda", 0)
for (b, m, r) in [(168, 2736, 50), (96, 1024, 50), (2, 4096, 64), (2, 4096, 32), (1, 16384, 64)]:
    X = torch.randn(b, m, 64, device=dev)
    for _ in range(3):
        Q = ops.thin_qr(X, r)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        Q = ops.thin_qr(X, r)
    e1.record()
    torch.cuda.synchronize()
    err = float((Q.transpose(1, 2) @ Q - torch.eye(r, device=dev)).abs().max())
    print(f"b={b} m={m} r={r}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us  |QtQ-I|={err:.2e}")
