"""SoW linear layer on sm_100a kernels -- host-side mirror of ``tn_gradient.layer.sow`` (reference file
tn_gradient/layer/sow.py).

Same constructor, attributes, parameter names and state-dict keys as the reference:
``acc_downweight`` / ``acc_upweight`` (frozen accumulation, empty until the first merge), ``downscale_weights[i]``
(in, r), ``upscale_weights[i]`` (r, out), ``bias``.  The math is not PyTorch: ``forward`` dispatches to one
autograd Function whose forward/backward call the C ABI (fused tcgen05 GEMMs), ``accumulate`` to the grouped
merge kernel.  CUDA only -- CPU tensors raise (no fallback).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops
from ._lib import SowB200Error


@dataclass
class SoWArgs:
    """Legacy argument bundle that scripts/run_glue.py:54,564 imports from ``tn_gradient.layer.sow`` (it only
    survives as a commented-out dataclass in the reference, tn_gradient/prepare.py:16-25)."""
    device: Optional[str] = None
    dtype: Optional[torch.dtype] = None
    init_method: str = "normal_QR"
    rank: int = 16
    n_iter: int = 5
    scale: float = 1


class SoWParameter(nn.ParameterList):
    """``n_iter`` factor matrices of one shape; ``from_weights`` swaps ``.data`` so that Parameter identity (held by
    optimizers and DDP) survives a merge.  Mirrors tn_gradient/layer/sow.py:15-42."""

    def __init__(self, in_features: int, out_features: int, n_iter: int = 1, device=None, dtype=None) -> None:
        super().__init__(
            [nn.Parameter(torch.empty(in_features, out_features, device=device, dtype=dtype)) for _ in range(n_iter)]
        )
        self.in_features = in_features
        self.out_features = out_features
        self.n_iter = n_iter

    def from_weights(self, weights: Sequence[torch.Tensor]) -> None:
        for i, w in enumerate(weights):
            self[i].data = w.data

    def extra_repr(self) -> str:
        return f"{self.n_iter} x ({self.in_features}, {self.out_features})"


# ---------------------------------------------------------------------------------------------------------
# autograd bridge
# ---------------------------------------------------------------------------------------------------------

def _bf16c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.bfloat16:
        t = t.to(torch.bfloat16)
    return t if t.is_contiguous() else t.contiguous()


class _SoWLinearFn(torch.autograd.Function):
    """y = x.W + scale*(x.A).B + bias with W frozen.  Saves x and t = scale*x.A (T x r_pad), never forms dW.

    Inputs of other dtypes follow the bf16 compute policy: cast to bf16 on the way in, results cast
    back to the caller's dtype.  ``W_c`` is the bf16 compute copy of W (W itself when W is bf16)."""

    @staticmethod
    def forward(ctx, x, W_c, A, B, bias, scale):
        out_dtype = x.dtype
        fin = A.shape[0]
        lead = x.shape[:-1]
        if x.numel() == 0:                      # empty batch: nothing to launch (the reference returns an empty tensor too)
            ctx.empty = True
            ctx.meta = (x.dtype, A, B, bias, lead, fin)
            return x.new_zeros(*lead, B.shape[1])
        ctx.empty = False
        x2 = _bf16c(x.reshape(-1, fin))
        A_c, B_c, bias_c = _bf16c(A), _bf16c(B), _bf16c(bias)
        y, t = ops.linear_fwd(x2, W_c, A_c, B_c, bias_c, scale)
        ctx.save_for_backward(x2, t, W_c, A_c, B_c)
        ctx.scale = float(scale)
        ctx.meta = (out_dtype, A.dtype, B.dtype, None if bias is None else bias.dtype, lead, fin)
        y = y.reshape(*lead, B.shape[1])
        return y if out_dtype == torch.bfloat16 else y.to(out_dtype)

    @staticmethod
    def backward(ctx, dy):
        if ctx.empty:
            _, A, B, bias, lead, fin = ctx.meta
            need_x, _, need_A, need_B, need_bias, _ = ctx.needs_input_grad
            return (dy.new_zeros(*lead, fin) if need_x else None, None, torch.zeros_like(A) if need_A else None,
                    torch.zeros_like(B) if need_B else None,
                    torch.zeros_like(bias) if (need_bias and bias is not None) else None, None)
        x2, t, W_c, A_c, B_c = ctx.saved_tensors
        out_dtype, a_dt, b_dt, bias_dt, lead, fin = ctx.meta
        need_x, _, need_A, need_B, need_bias, _ = ctx.needs_input_grad
        dy2 = _bf16c(dy.reshape(-1, dy.shape[-1]))
        dt, dA, dB, dbias = ops.linear_bwd_factors(dy2, x2, t, B_c, ctx.scale, bool(need_bias), fin)
        dx = None
        if need_x:
            dx = ops.linear_bwd_dx(dy2, dt, W_c, A_c).reshape(*lead, fin)
            if out_dtype != torch.bfloat16:
                dx = dx.to(out_dtype)
        dA = dA.to(a_dt) if need_A else None
        dB = dB.to(b_dt) if need_B else None
        dbias = dbias.to(bias_dt) if need_bias else None
        return dx, None, dA, dB, dbias, None


@torch.compiler.disable
def sow_linear(x, W_c, A, B, bias, scale):
    """Kernel-backed SoW linear.  Opaque to torch.compile (scripts/finetune.py:486-487 compiles the model): the C-ABI
    calls are not traceable, so dynamo breaks the graph here and runs this call eagerly."""
    return _SoWLinearFn.apply(x, W_c, A, B, bias, scale)


# ---------------------------------------------------------------------------------------------------------
# QR initialisation helpers (thin-QR kernel instead of a full in x out QR)
# ---------------------------------------------------------------------------------------------------------

def _cuda_device_for(t: torch.Tensor) -> torch.device:
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise SowB200Error(
            'init_method="normal_QR" runs its QR on the GPU (the reference hard-codes .to("cuda") too, '
            "tn_gradient/layer/sow.py:91) and no CUDA device is available"
        )
    return torch.device("cuda", torch.cuda.current_device())


def qr_init_factors(in_features: int, out_features: int, rank: int, device: torch.device, want_b: bool,
                    std: float = 0.02, generator: Optional[torch.Generator] = None):
    """A = first ``rank`` columns of Q of QR(G), G ~ N(0, std) (in x out); B = A^T G (= R[:rank]) if requested.

    Q[:, :rank] of a Householder QR depends on the first ``rank`` columns only, so the thin-QR
    kernel on those columns yields the same subspace at 1/(out/rank) of the cost; columns are sign-normalised
    (diag(R) >= 0) whereas LAPACK's signs vary, which leaves A.B and the distribution of A unchanged.
    """
    if want_b:
        G = torch.empty((in_features, out_features), dtype=torch.float32, device=device).normal_(0.0, std, generator=generator)
        A = ops.thin_qr(G, rank)
        B = ops.project(G, A)
        return A, B
    G = torch.empty((in_features, rank), dtype=torch.float32, device=device).normal_(0.0, std, generator=generator)
    return ops.thin_qr(G, rank), None


class SoWLinear(nn.Module):
    """Drop-in for tn_gradient.layer.sow.SoWLinear (tn_gradient/layer/sow.py:45-181)."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, rank: int = 16, n_iter: int = 1,
                 scale: float = 1, init_method: str = "normal_QR", device=None, dtype=None, init_params=True) -> None:
        super().__init__()
        fk = {"device": device, "dtype": dtype}
        self.in_features = in_features
        self.out_features = out_features
        self.n_iter = n_iter
        self.rank = rank
        self.scale = scale
        self.virtual_rank = min(rank * n_iter, in_features, out_features)
        self.init_method = init_method

        self.acc_upweight = nn.Parameter(torch.empty(0), requires_grad=False)
        self.acc_downweight = nn.Parameter(torch.empty(0), requires_grad=False)
        self.downscale_weights = SoWParameter(in_features, rank, n_iter=n_iter, device=device, dtype=dtype)
        self.upscale_weights = SoWParameter(rank, out_features, n_iter=n_iter, device=device, dtype=dtype)
        if bias:
            self.bias = nn.Parameter(torch.empty(out_features, **fk))
        else:
            self.register_parameter("bias", None)
        self._w_shadow = None       # bf16 compute copy of a non-bf16 acc_downweight
        self._w_shadow_key = None
        if init_params:
            self.reset_parameters()

    # ---- initialisation (sow.py:89-105) ------------------------------------------------------------------
    def reset_parameters(self, reset_scale=1.0) -> None:
        for i in range(self.n_iter):
            if i / self.n_iter >= 1 - reset_scale:
                A_p, B_p = self.downscale_weights[i], self.upscale_weights[i]
                if self.init_method == "normal_QR":
                    dev = _cuda_device_for(A_p)
                    A, B = qr_init_factors(self.in_features, self.out_features, self.rank, dev, want_b=True)
                    with torch.no_grad():
                        A_p.copy_(A.to(A_p.device))
                        B_p.copy_(B.to(B_p.device))
                else:
                    nn.init.normal_(A_p, std=0.02)
                    nn.init.normal_(B_p, std=0.02)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    # ---- forward (sow.py:107-126) ---------------------------------------------------------------------------
    def _compute_weight(self) -> Optional[torch.Tensor]:
        W = self.acc_downweight
        if W.numel() == 0:
            return None
        if W.dtype == torch.bfloat16 and W.is_contiguous():
            return W
        key = (W.data_ptr(), W._version, W.dtype)
        if self._w_shadow is None or self._w_shadow_key != key:
            self._w_shadow = W.detach().to(torch.bfloat16).contiguous()
            self._w_shadow_key = key
        return self._w_shadow

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise SowB200Error("SoWLinear.forward needs CUDA tensors (sm_100a kernels only, no CPU fallback)")
        A_list = list(self.downscale_weights)
        B_list = list(self.upscale_weights)
        if len(A_list) == 1:
            A, B = A_list[0], B_list[0]
        else:
            # sum_i (x.A_i).B_i == (x.[A_1|..|A_n]).[B_1;..;B_n]: one fused call; autograd splits the grads
            A = torch.cat(A_list, dim=1)
            B = torch.cat(B_list, dim=0)
        factored = self.acc_downweight.numel() != 0 and self.acc_upweight.numel() != 0
        W_c = None if factored else self._compute_weight()
        out = sow_linear(x, W_c, A, B, None if factored else self.bias, self.scale)
        if factored:
            # compat branch (direct construction with virtual_rank < min(in,out); never reached through
            # prepare_sow, prepare.py:120): the frozen factored accumulation goes through cuBLAS
            out = out + (x @ self.acc_downweight.to(x.dtype)) @ self.acc_upweight.to(x.dtype)
            if self.bias is not None:
                out = out + self.bias.to(out.dtype)
        return out

    # ---- merge (sow.py:128-178) ---------------------------------------------------------------------------
    def accumulate(self) -> None:
        accumulate_modules([self])

    def extra_repr(self) -> str:
        return (f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}, "
                f"rank={self.rank}, n_iter={self.n_iter}")


# ---------------------------------------------------------------------------------------------------------
# grouped merge + re-initialisation over many modules
# ---------------------------------------------------------------------------------------------------------

def _merge_dense(mods: List[SoWLinear]) -> None:
    """Dense branch (sow.py:151-153) for all modules in one grouped launch per rank chunk."""
    items = []
    post = []
    for mod in mods:
        A_list = [a.detach() for a in mod.downscale_weights]
        B_list = [b.detach() for b in mod.upscale_weights]
        dev = A_list[0].device
        if not dev.type == "cuda":
            raise SowB200Error("SoWLinear.accumulate needs the module on a CUDA device (no CPU fallback)")
        A = A_list[0] if len(A_list) == 1 else torch.cat(A_list, dim=1)
        B = B_list[0] if len(B_list) == 1 else torch.cat(B_list, dim=0)
        pdtype = A.dtype
        A_c, B_c = _bf16c(A), _bf16c(B)
        W_old = mod.acc_downweight
        expanded = W_old.numel() != 0 and mod.acc_upweight.numel() != 0
        if expanded:
            # last QR-growth step reached full rank: expand the factored accumulation once (sow.py:137-138); the
            # product is a temporary, so the merged result must be bound as the new acc_downweight below
            tgt = W_old.dtype
            W_tmp = _bf16c((W_old.detach() @ mod.acc_upweight.detach()).to(dev))
            items.append((W_tmp, W_tmp, A_c, B_c, mod.scale))
            post.append((mod, W_tmp, tgt))
            continue
        has_prev = W_old.numel() != 0
        if has_prev and W_old.dtype == torch.bfloat16 and W_old.is_contiguous() and W_old.device == dev:
            items.append((W_old.data, W_old.data, A_c, B_c, mod.scale))      # in-place RMW: pointer stays stable
            post.append((mod, None, None))
        else:
            W_new = torch.empty((mod.in_features, mod.out_features), dtype=torch.bfloat16, device=dev)
            prev = _bf16c(W_old.detach().to(dev)) if has_prev else None
            items.append((W_new, prev, A_c, B_c, mod.scale))
            post.append((mod, W_new, W_old.dtype if has_prev else pdtype))
    ops.merge_grouped(items)
    for mod, W_new, tgt_dtype in post:
        if W_new is not None:
            W_final = W_new if tgt_dtype == torch.bfloat16 else W_new.to(tgt_dtype)
            mod.acc_downweight = nn.Parameter(W_final, requires_grad=False)
            mod.acc_upweight = nn.Parameter(torch.empty(0, device=W_final.device), requires_grad=False)
        mod._w_shadow = None
        mod._w_shadow_key = None


def _merge_factored(mod: SoWLinear) -> None:
    """Factored / QR-growth branch (sow.py:137-150), only reachable by direct construction with
    virtual_rank < min(in, out).  One-off compat path: cuBLAS + cuSOLVER through torch."""
    from .utils import qr_weight
    acc = None
    for a, b in zip(mod.downscale_weights, mod.upscale_weights):
        term = a.detach() @ b.detach()
        acc = term if acc is None else acc + term
    acc = mod.scale * acc
    if mod.acc_downweight.numel() != 0 and mod.acc_upweight.numel() != 0:
        acc = acc + mod.acc_downweight @ mod.acc_upweight
    elif mod.acc_downweight.numel() != 0:
        acc = acc + mod.acc_downweight
    Q, R = qr_weight(acc, rank=mod.virtual_rank)
    mod.acc_downweight = nn.Parameter(Q.contiguous(), requires_grad=False)
    mod.acc_upweight = nn.Parameter(R.contiguous(), requires_grad=False)
    mod.virtual_rank = min(mod.virtual_rank + mod.rank * mod.n_iter, mod.in_features, mod.out_features)


def _reinit(mods: List[SoWLinear], sync: bool) -> None:
    """A <- QR-init / N(0, .02), B <- 0 (sow.py:157-178); .data swap keeps Parameter identity.

    With torch.distributed initialised and ``sync`` the new A is broadcast from rank 0, which makes the replicas
    consistent by construction (the reference relies on identical RNG state on every rank)."""
    import torch.distributed as dist
    do_sync = sync and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    # batch the thin QRs of all normal_QR modules that share (in, rank, device)
    groups = {}
    for mod in mods:
        if mod.init_method == "normal_QR":
            dev = mod.downscale_weights[0].device
            groups.setdefault((mod.in_features, mod.rank, dev), []).append(mod)
    new_A = {}
    for (fin, r, dev), ms in groups.items():
        n = sum(m.n_iter for m in ms)
        # the reference draws in the accumulation dtype (bf16-rounded Gaussian, sow.py:163-165,170)
        wdt = ms[0].acc_downweight.dtype if ms[0].acc_downweight.numel() else ms[0].downscale_weights[0].dtype
        G = torch.empty((n, fin, r), dtype=torch.float32, device=dev).normal_(0.0, 0.02)
        if wdt != torch.float32:
            G = G.to(wdt).to(torch.float32)
        Q = ops.thin_qr(G, r)
        k = 0
        for m in ms:
            for i in range(m.n_iter):
                new_A[(id(m), i)] = Q[k]
                k += 1
    for mod in mods:
        downs, ups = [], []
        for i, (a, b) in enumerate(zip(mod.downscale_weights, mod.upscale_weights)):
            if mod.init_method == "normal_QR":
                a_new = new_A[(id(mod), i)].to(a.dtype)
            else:
                a_new = torch.empty_like(a).normal_(std=0.02)
            if do_sync:
                dist.broadcast(a_new, src=0)
            downs.append(a_new.contiguous())
            ups.append(torch.zeros_like(b))
        mod.downscale_weights.from_weights(downs)
        mod.upscale_weights.from_weights(ups)


@torch.no_grad()
def accumulate_modules(mods: Sequence[SoWLinear], sync_reinit: bool = True) -> None:
    """SoWLinear.accumulate over a list of modules: grouped merge, then factor re-initialisation."""
    mods = list(mods)
    for m in mods:
        if not m.downscale_weights[0].is_cuda:
            raise SowB200Error("SoWLinear.accumulate needs the module on a CUDA device (sm_100a kernels only, no CPU fallback)")
    dense = [m for m in mods if m.virtual_rank >= min(m.in_features, m.out_features)]
    fact = [m for m in mods if m.virtual_rank < min(m.in_features, m.out_features)]
    by_dev = {}
    for m in dense:
        by_dev.setdefault(m.downscale_weights[0].device, []).append(m)
    for ms in by_dev.values():
        _merge_dense(ms)
    for m in fact:
        _merge_factored(m)
    _reinit(mods, sync_reinit)
