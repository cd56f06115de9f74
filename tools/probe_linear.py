"""GPU bring-up probe for the tcgen05 GEMM variants behind sow_linear_{fwd,bwd_*} (debug aid, not a parity test).

Calls the C ABI through ctypes on raw torch device pointers and compares every output against fp32 torch math,
printing an error map per output so descriptor/layout bugs can be localised from one run.
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sow_b200 import _lib  # noqa: E402


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def report(name, got, ref, blk_r=16, blk_c=32):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    rel = (got - ref).norm() / (ref.norm() + 1e-30)
    ok = rel < 1e-2
    print(f"[{'OK ' if ok else 'BAD'}] {name:10s} shape={tuple(got.shape)} rel={rel:.3e} maxabs={err.max():.3e} "
          f"ref_rms={ref.pow(2).mean().sqrt():.3e} nan={int(torch.isnan(got).sum())}")
    if not ok:
        R, C = got.shape
        rr, cc = min(R, 128), min(C, 256)
        sub = err[:rr, :cc]
        nbr, nbc = (rr + blk_r - 1) // blk_r, (cc + blk_c - 1) // blk_c
        print(f"   error map of first {rr}x{cc} block ({blk_r}x{blk_c} cells, max abs err; ref rms above):")
        for i in range(nbr):
            row = []
            for j in range(nbc):
                cell = sub[i * blk_r:(i + 1) * blk_r, j * blk_c:(j + 1) * blk_c]
                row.append(f"{cell.max():8.2e}")
            print("   " + " ".join(row))
        # which whole rows / cols are good?
        row_bad = (err.max(dim=1).values > 0.05 * ref.abs().max()).nonzero().flatten()
        col_bad = (err.max(dim=0).values > 0.05 * ref.abs().max()).nonzero().flatten()
        print(f"   bad rows: {row_bad.numel()}/{R} first {row_bad[:16].tolist()} ; bad cols: {col_bad.numel()}/{C} first {col_bad[:16].tolist()}")
    return ok


def run_case(T, fin, fout, r, scale, with_W, with_bias, seed=0):
    lib = _lib.load()
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(seed)
    bf = torch.bfloat16
    x = torch.randn(T, fin, generator=g).to(dev, bf)
    W = (torch.randn(fin, fout, generator=g) * 0.02).to(dev, bf) if with_W else None
    A = (torch.randn(fin, r, generator=g) * 0.05).to(dev, bf)
    B = (torch.randn(r, fout, generator=g) * 0.05).to(dev, bf)
    bias = torch.randn(fout, generator=g).to(dev, bf) if with_bias else None
    dy = torch.randn(T, fout, generator=g).to(dev, bf)
    r_pad = lib.sow_rank_pad(r)
    print(f"--- T={T} in={fin} out={fout} r={r} r_pad={r_pad} scale={scale} W={with_W} bias={with_bias}")

    y = torch.full((T, fout), float("nan"), device=dev, dtype=bf)
    t = torch.full((T, r_pad), float("nan"), device=dev, dtype=bf)
    ws_f = lib.sow_workspace_bytes(_lib.OP_LINEAR_FWD, T, fin, fout, r)
    ws_b = lib.sow_workspace_bytes(_lib.OP_LINEAR_BWD, T, fin, fout, r)
    ws = torch.empty(max(ws_f, ws_b), device=dev, dtype=torch.uint8)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    rc = lib.sow_linear_fwd(ptr(x), ptr(W), ptr(A), ptr(B), ptr(bias), ptr(y), ptr(t), T, fin, fout, r, scale,
                            _lib.SOWB_BF16, ptr(ws), ws.numel(), stream)
    _lib.check(rc, "sow_linear_fwd")
    torch.cuda.synchronize()
    xf, Af, Bf = x.float(), A.float(), B.float()
    t_ref = scale * (xf @ Af)
    ok = True
    ok &= report("t", t[:, :r], t_ref)
    ok &= report("t_pad", t[:, r:], torch.zeros(T, r_pad - r, device=dev)) if r_pad > r else True
    y_ref = t[:, :r].float() @ Bf
    if with_W:
        y_ref = y_ref + xf @ W.float()
    if with_bias:
        y_ref = y_ref + bias.float()
    ok &= report("y", y, y_ref)

    dt = torch.full((T, r_pad), float("nan"), device=dev, dtype=bf)
    dA = torch.full((fin, r), float("nan"), device=dev, dtype=bf)
    dB = torch.full((r, fout), float("nan"), device=dev, dtype=bf)
    dbias = torch.full((fout,), float("nan"), device=dev, dtype=bf) if with_bias else None
    rc = lib.sow_linear_bwd_factors(ptr(dy), ptr(x), ptr(t), ptr(B), ptr(dt), ptr(dA), ptr(dB), ptr(dbias), T, fin,
                                    fout, r, scale, _lib.SOWB_BF16, ptr(ws), ws.numel(), stream)
    _lib.check(rc, "sow_linear_bwd_factors")
    torch.cuda.synchronize()
    dyf = dy.float()
    dt_ref = scale * (dyf @ Bf.T)
    ok &= report("dt", dt[:, :r], dt_ref)
    ok &= report("dB", dB, t[:, :r].float().T @ dyf)
    ok &= report("dA", dA, xf.T @ dt[:, :r].float())
    if with_bias:
        ok &= report("dbias", dbias[None, :], dyf.sum(0)[None, :])

    dx = torch.full((T, fin), float("nan"), device=dev, dtype=bf)
    rc = lib.sow_linear_bwd_dx(ptr(dy), ptr(dt), ptr(W), ptr(A), ptr(dx), T, fin, fout, r, _lib.SOWB_BF16, ptr(ws),
                               ws.numel(), stream)
    _lib.check(rc, "sow_linear_bwd_dx")
    torch.cuda.synchronize()
    dx_ref = dt[:, :r].float() @ Af.T
    if with_W:
        dx_ref = dx_ref + dyf @ W.float().T
    ok &= report("dx", dx, dx_ref)
    return ok


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    all_ok = True
    all_ok &= run_case(256, 128, 256, 50, 1.0, False, False)
    all_ok &= run_case(256, 128, 256, 50, 0.5, True, False)
    all_ok &= run_case(384, 192, 328, 8, 0.125, True, True)        # ragged tails everywhere
    all_ok &= run_case(4096, 1024, 2736, 50, 1.0, True, False)     # llama_350m gate/up
    all_ok &= run_case(4096, 2736, 1024, 50, 1.0, True, False)     # llama_350m down
    print("ALL OK" if all_ok else "SOME BAD")
    sys.exit(0 if all_ok else 1)
