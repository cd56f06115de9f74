"""GPU bring-up probe for merge / thin-QR / TT / Adam kernels (debug aid; parity tests live in tests/)."""
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sow_b200 import ops  # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-300))


def status(name, err, tol):
    print(f"[{'OK ' if err < tol else 'BAD'}] {name}: {err:.3e} (tol {tol:g})", flush=True)
    return err < tol


def probe_merge():
    ok = True
    dev = "cuda"
    torch.manual_seed(0)
    items, refs = [], []
    for (fin, fout, r, s, prev) in [(128, 128, 50, 1.0, True), (1024, 2736, 50, 0.5, True), (2736, 1024, 50, 1.0, False),
                                    (192, 328, 8, 0.125, True), (256, 512, 200, 1.0, True), (1000, 1024, 70, 1.0, False)]:
        A = (torch.randn(fin, r, device=dev) * 0.05).bfloat16()
        B = (torch.randn(r, fout, device=dev) * 0.05).bfloat16()
        Wp = (torch.randn(fin, fout, device=dev) * 0.02).bfloat16() if prev else None
        ref = s * (A.float() @ B.float()) + (Wp.float() if prev else 0)
        W = Wp if prev else torch.full((fin, fout), float("nan"), device=dev, dtype=torch.bfloat16)
        items.append((W, Wp, A, B, s))
        refs.append(ref)
    ops.merge_grouped(items)
    torch.cuda.synchronize()
    for (W, _, A, B, s), ref in zip(items, refs):
        ok &= status(f"merge {tuple(W.shape)} r={A.shape[1]}", rel(W, ref), 4e-3)
    return ok


def probe_qr_project():
    ok = True
    dev = "cuda"
    torch.manual_seed(1)
    for (b, m, n, r) in [(1, 200, 300, 16), (3, 1024, 64, 50), (1, 4096, 4096, 64), (2, 256, 2048, 8), (1, 729, 9, 4),
                         (1, 6720, 1000, 64), (1, 300, 80, 70)]:
        X = torch.randn(b, m, n, device=dev)
        t0 = time.time()
        Q = ops.thin_qr(X, r)
        torch.cuda.synchronize()
        dt = time.time() - t0
        eye = torch.eye(r, device=dev).expand(b, r, r)
        orth = float((Q.transpose(1, 2) @ Q - eye).abs().max())
        Qref, _ = torch.linalg.qr(X[:, :, :r].double().cpu())
        Qd = Q.double().cpu()
        proj = float(((Qd @ Qd.transpose(1, 2)) @ X[:, :, :r].double().cpu() - X[:, :, :r].double().cpu()).abs().max())
        sign = torch.sign(torch.diagonal(Qref.transpose(1, 2) @ Qd, dim1=1, dim2=2))
        qerr = rel(Qd, Qref * sign[:, None, :])
        ok &= status(f"thin_qr b={b} m={m} r={r} orth ({dt*1e3:.1f} ms incl sync)", orth, 5e-6)
        ok &= status("   span residual", proj, 1e-4)
        ok &= status("   vs LAPACK Q (sign-fixed)", qerr, 1e-4)
        R = ops.project(X, Q)
        torch.cuda.synchronize()
        Rref = Q.double().transpose(1, 2) @ X.double()
        ok &= status(f"   project R=Q^T L n={n}", rel(R, Rref), 2e-6)
    return ok


def probe_interleave_fused():
    ok = True
    dev = "cuda"
    torch.manual_seed(2)
    for (M, N, order, dtype) in [(50, 30, 2, torch.float32), (64, 48, 2, torch.bfloat16), (81, 81, 4, torch.float32),
                                 (64, 27, 3, torch.float32), (300, 200, 2, torch.float32)]:
        mm = math.ceil(M ** (1 / order))
        nn = math.ceil(N ** (1 / order))
        src = torch.randn(M, N, device=dev).to(dtype)
        out = ops.interleave(src, mm, nn, order)
        pad = torch.zeros(mm ** order, nn ** order, device=dev)
        pad[:M, :N] = src.float()
        t = pad.reshape((mm,) * order + (nn,) * order)
        perm = [i for pair in zip(range(order), range(order, 2 * order)) for i in pair]
        ref = t.permute(*perm).contiguous().flatten()
        ok &= status(f"interleave {M}x{N} order {order} {dtype}", float((out - ref).abs().max()), 1e-30)
        back = ops.deinterleave(out, M, N, mm, nn, order, dtype)
        ok &= status("   deinterleave roundtrip", float((back.float() - src.float()).abs().max()), 1e-30)
    # fused order-2 adam
    for (M, N, r, dtype, wd) in [(50, 30, 6, torch.float32, 0.0), (64, 48, 8, torch.float32, 0.01),
                                 (300, 200, 64, torch.float32, 0.0), (128, 96, 16, torch.bfloat16, 0.0)]:
        mm = math.ceil(M ** 0.5)
        nn = math.ceil(N ** 0.5)
        P = mm * nn
        G1m, G2m = torch.randn(P, r, device=dev) * 0.3, torch.randn(r, P, device=dev) * 0.3
        G1v, G2v = torch.randn(P, r, device=dev) * 0.3, torch.randn(r, P, device=dev) * 0.3
        p0 = (torch.randn(M, N, device=dev) * 0.1).to(dtype)
        g = torch.randn(M, N, device=dev).to(dtype)
        b1, b2, eps, ss, lr = 0.9, 0.999, 1e-8, 3e-3, 1e-2
        mprev = ops.deinterleave((G1m @ G2m).flatten(), M, N, mm, nn, 2)
        vprev = ops.deinterleave((G1v @ G2v).flatten(), M, N, mm, nn, 2).clamp_min(0)
        gf, pf = g.double(), p0.double()
        mref = b1 * mprev.double() + (1 - b1) * gf
        vref = b2 * vprev.double() + (1 - b2) * gf * gf
        pref = pf - ss * mref / (vref.sqrt() + eps)
        if wd > 0:
            pref = pref - lr * wd * pref
        p = p0.clone()
        m_out, v_out = ops.tt_adam_fused2(p, g, (G1m, G2m), (G1v, G2v), mm, nn, b1, b2, eps, ss, lr * wd, False)
        torch.cuda.synchronize()
        tol = 4e-3 if dtype == torch.bfloat16 else 2e-6
        ok &= status(f"fused2 p {M}x{N} r={r} {dtype}", rel(p, pref), tol)
        ok &= status("   m_out", rel(ops.deinterleave(m_out.flatten(), M, N, mm, nn, 2), mref), 2e-6)
        ok &= status("   v_out", rel(ops.deinterleave(v_out.flatten(), M, N, mm, nn, 2), vref), 2e-6)
        pad_mask = ops.interleave(torch.ones(M, N, device=dev), mm, nn, 2) == 0
        ok &= status("   padding is zero", float(m_out.flatten()[pad_mask].abs().max()) if pad_mask.any() else 0.0, 1e-30)
    A = torch.randn(500, 24, device=dev)
    B = torch.randn(24, 700, device=dev)
    ok &= status("matmul_rk", rel(ops.matmul_rk(A, B), A.double() @ B.double()), 2e-6)
    A = torch.randn(130, 100, device=dev)
    B = torch.randn(100, 70, device=dev)
    ok &= status("matmul_rk r=100", rel(ops.matmul_rk(A, B), A.double() @ B.double()), 2e-6)
    return ok


def probe_adam():
    ok = True
    dev = "cuda"
    torch.manual_seed(3)
    for dtype in (torch.bfloat16, torch.float32):
        shapes = [(1024, 50), (50, 2736), (7,), (100001,), (32000, 64)]
        ps = [(torch.randn(*s, device=dev) * 0.1).to(dtype) for s in shapes]
        gs = [torch.randn(*s, device=dev).to(dtype) for s in shapes]
        ms = [(torch.randn(*s, device=dev) * 0.01).to(dtype) for s in shapes]
        vs = [(torch.rand(*s, device=dev) * 0.01).to(dtype) for s in shapes]
        lr, b1, b2, eps, wd, step = 1e-2, 0.9, 0.999, 1e-8, 0.1, 3
        refs = []
        for p, g, m, v in zip(ps, gs, ms, vs):
            pd, gd, md, vd = p.double(), g.double(), m.double(), v.double()
            pd = pd * (1 - lr * wd)
            md = b1 * md + (1 - b1) * gd
            vd = b2 * vd + (1 - b2) * gd * gd
            bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
            pd = pd - (lr / bc1) * md / (vd.sqrt() / math.sqrt(bc2) + eps)
            refs.append((pd, md, vd))
        chunks = ops.build_adam_chunks(ps, gs, ms, vs)
        ops.adam_multi(chunks, dtype, lr, b1, b2, eps, wd, 1 - b1 ** step, 1 - b2 ** step, True)
        torch.cuda.synchronize()
        tol = 4e-3 if dtype == torch.bfloat16 else 1e-6
        for (p, m, v), (pr, mr, vr), s in zip(zip(ps, ms, vs), refs, shapes):
            ok &= status(f"adam {dtype} {s}", max(rel(p, pr), rel(m, mr), rel(v, vr)), tol)
    return ok


if __name__ == "__main__":
    which = sys.argv[1:] or ["merge", "qr", "tt", "adam"]
    ok = True
    for w in which:
        fn = {"merge": probe_merge, "qr": probe_qr_project, "tt": probe_interleave_fused, "adam": probe_adam}[w]
        try:
            ok &= fn()
        except Exception as e:  # keep going so one run reports on every kernel
            import traceback
            traceback.print_exc()
            ok = False
    print("ALL OK" if ok else "SOME BAD")
    sys.exit(0 if ok else 1)
