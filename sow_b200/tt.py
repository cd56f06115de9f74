"""Tensor-train container -- host-side mirror of ``tn_gradient.tt.TensorTrain`` (reference file tn_gradient/tt.py).

Hot-path methods run on the CUDA kernels and require CUDA tensors (no CPU fallback):
    from_matrix / from_tensor / decompose   -> tt_interleave + sow_thin_qr + tt_project      (tt.py:26-67,111-140)
    reconstruct / to_tensor / to_matrix     -> tt_matmul_rk chain + tt_deinterleave          (tt.py:213-247)
The decomposition is the same truncated-QR projection as the reference's complete-QR sweep; bases differ by an
orthogonal gauge (sign of each column), so compare reconstructions, never cores.

The remaining algebra (+, *, scalar *, add_, round, orthogonalize, inner/norm, sqrt/sqrtinv, reciprocal) is not on
the hot path and stays in PyTorch on whatever device the cores live ("compat").
"""
from __future__ import annotations

from math import ceil, floor, log
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import SowB200Error

_LETTERS = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


def _uniform(shape) -> bool:
    return len(set(int(s) for s in shape)) == 1


class _Contraction:
    """Cached contraction handle stored in ``TensorTrain.contract_expr`` (the reference caches an opt_einsum
    expression there and TTAdam copies it into its state, ttadam.py:73,81)."""

    def __init__(self, shapes):
        self.shapes = [tuple(s) for s in shapes]

    def __call__(self, *cores):
        return _reconstruct_cores(list(cores))


def reconstruct_interleaved(cores: Sequence[torch.Tensor]) -> torch.Tensor:
    """Flat fp32 tensor in (i1,o1,...,id,od) order from cores (r_k, i_k, o_k, r_{k+1}): the chain of small-K matmuls
    (tt.py:213-237) already produces this layout, which is also what the decomposition sweep consumes."""
    c0 = cores[0]
    if not c0.is_cuda:
        raise SowB200Error("TensorTrain.reconstruct needs CUDA cores (sm_100a kernels only, no CPU fallback)")
    res = c0.detach().to(torch.float32).reshape(-1, c0.shape[-1])
    for c in cores[1:]:
        cm = c.detach().to(torch.float32).reshape(c.shape[0], -1)
        res = ops.matmul_rk(res, cm).reshape(-1, c.shape[-1])
    return res.reshape(-1)


def _reconstruct_cores(cores: Sequence[torch.Tensor]) -> torch.Tensor:
    """(i1..id, o1..od)-shaped fp32 tensor from cores (r_k, i_k, o_k, r_{k+1}) via the small-K matmul kernel."""
    c0 = cores[0]
    if not c0.is_cuda:
        raise SowB200Error("TensorTrain.reconstruct needs CUDA cores (sm_100a kernels only, no CPU fallback)")
    d = len(cores)
    ins = [int(c.shape[1]) for c in cores]
    outs = [int(c.shape[2]) for c in cores]
    if cores[0].shape[0] != 1 or cores[-1].shape[-1] != 1:
        raise ValueError("boundary TT ranks must be 1")
    res = c0.detach().to(torch.float32).reshape(-1, c0.shape[-1])
    for c in cores[1:]:
        cm = c.detach().to(torch.float32).reshape(c.shape[0], -1)
        res = ops.matmul_rk(res, cm).reshape(-1, c.shape[-1])
    flat = res.reshape(-1)
    if _uniform(ins) and _uniform(outs):
        mm, nn_ = ins[0], outs[0]
        mat = ops.deinterleave(flat, mm ** d, nn_ ** d, mm, nn_, d, torch.float32)
        return mat.reshape(tuple(ins) + tuple(outs))
    inter = [s for pair in zip(ins, outs) for s in pair]
    perm = list(range(0, 2 * d, 2)) + list(range(1, 2 * d, 2))
    return flat.reshape(inter).permute(*perm).contiguous()


class TensorTrain:

    def __init__(self, ranks, input_shape, output_shape, device=None) -> None:
        self.order = len(ranks) - 1
        self.ranks = ranks                      # aliased, like the reference (tt.py:17)
        self.input_shape = input_shape
        self.output_shape = output_shape
        self.cores: List[Optional[torch.Tensor]] = [None for _ in range(self.order)]
        self.device = device
        self.contract_expr = None

    # ---- constructors ------------------------------------------------------------------------------------
    @staticmethod
    def from_tensor(tensor: torch.Tensor, ranks: list):
        """tensor axes are (*input_shape, *output_shape) (tt.py:26-35)."""
        d = tensor.dim() // 2
        tt = TensorTrain(ranks, tuple(tensor.shape[:d]), tuple(tensor.shape[d:]))
        perm = [i for pair in zip(range(d), range(d, 2 * d)) for i in pair]
        tt.decompose(tensor.permute(*perm))
        return tt

    @staticmethod
    def from_cores(cores):
        tt = TensorTrain([c.shape[0] for c in cores] + [1], [c.shape[1] for c in cores], [c.shape[2] for c in cores],
                         device=cores[0].device)
        tt.cores = cores
        return tt

    @staticmethod
    def from_matrix(matrix: torch.Tensor, ranks: list, padding=True):
        """Pad to (mm^d, nn^d), fold, interleave, decompose (tt.py:48-67) -- pad + fold + interleave are one
        kernel (tt_interleave), no padded / permuted temporaries."""
        order = len(ranks) - 1
        M, N = matrix.shape
        mm = ceil(M ** (1 / order))
        nn_ = ceil(N ** (1 / order))
        if not padding and (mm ** order != M or nn_ ** order != N):
            raise RuntimeError(f"shape {tuple(matrix.shape)} cannot be folded into order {order} without padding")
        if not matrix.is_cuda:
            raise SowB200Error("TensorTrain.from_matrix needs a CUDA tensor (sm_100a kernels only, no CPU fallback)")
        src = matrix.detach()
        if src.dtype not in (torch.float32, torch.bfloat16):
            src = src.to(torch.float32)
        tt = TensorTrain(ranks, (mm,) * order, (nn_,) * order)
        if order == 2:
            # the unfolding is addressed through the index map: no padded / interleaved copy of the matrix at all
            r = ranks[1]
            if r > mm * nn_:
                raise RuntimeError(f"TT rank {r} exceeds the unfolding row count {mm * nn_} at core 0 "
                                   "(the reference raises a reshape error here, tt.py:135)")
            Q, R = ops.decompose2(src, mm, nn_, r)
            tt.cores = [Q.reshape(1, mm, nn_, r), R.reshape(r, mm, nn_, 1)]
            tt.device = matrix.device
            return tt
        cores = ops.decompose_nd(src, mm, nn_, ranks)          # interleave + sweep in one C-ABI call (inner ranks <= 64)
        if cores is not None:
            tt.cores = cores
            tt.device = matrix.device
            return tt
        flat = ops.interleave(src, mm, nn_, order)
        tt._decompose_interleaved(flat)
        return tt.to(matrix.device)

    @staticmethod
    def zeros(ranks, input_shape, output_shape, device="cpu"):
        tt = TensorTrain(ranks, input_shape, output_shape)
        tt.cores = [torch.zeros((ranks[i], input_shape[i], output_shape[i], ranks[i + 1])) for i in range(tt.order)]
        tt.to(device)
        return tt

    @staticmethod
    def ones(ranks, input_shape, output_shape, device="cpu"):
        tt = TensorTrain(ranks, input_shape, output_shape)
        tt.cores = [torch.ones((ranks[i], input_shape[i], output_shape[i], ranks[i + 1])) for i in range(tt.order)]
        tt.to(device)
        return tt

    # ---- bookkeeping -----------------------------------------------------------------------------------------
    def numel(self):
        return sum(core.numel() for core in self.cores)

    def to(self, device):
        self.device = device
        if self.cores[0].device == device:
            return self
        self.cores = [core.to(device) for core in self.cores]
        return self

    def clone(self):
        tt = TensorTrain(list(self.ranks), self.input_shape, self.output_shape)
        tt.cores = list(self.cores)
        tt.device = self.device
        return tt

    def detach(self):
        tt = TensorTrain(list(self.ranks), self.input_shape, self.output_shape)
        tt.cores = [core.detach() for core in self.cores]
        tt.device = self.device
        return tt

    def type(self, dtype):
        self.cores = [core.type(dtype) for core in self.cores]
        return self

    def requires_grad_(self, flag):
        for core in self.cores:
            core.requires_grad_(flag)
        return self

    def size(self):
        return [core.size() for core in self.cores]

    def to_params(self):
        cores = nn.ParameterList()
        for core in self.cores:
            cores.append(nn.Parameter(core))
        self.cores = cores
        return self

    # ---- decomposition (tt.py:111-140) -----------------------------------------------------------------
    def decompose(self, tensor: torch.Tensor):
        """``tensor`` is already interleaved (i1,o1,i2,o2,...)."""
        if not tensor.is_cuda:
            raise SowB200Error("TensorTrain.decompose needs a CUDA tensor (sm_100a kernels only, no CPU fallback)")
        flat = tensor.detach().to(torch.float32).contiguous().reshape(-1)
        return self._decompose_interleaved(flat)

    def _decompose_interleaved(self, flat: torch.Tensor):
        """Left-to-right sweep: Q_k = thin-QR(L_k[:, :r_{k+1}]), next L = Q_k^T L_k (projection kernel)."""
        cur = flat
        for k in range(self.order - 1):
            rows = self.ranks[k] * self.input_shape[k] * self.output_shape[k]
            L = cur.reshape(rows, -1)
            r = self.ranks[k + 1]
            if r > rows:
                raise RuntimeError(f"TT rank {r} exceeds the unfolding row count {rows} at core {k} "
                                   "(the reference raises a reshape error here, tt.py:135)")
            Q = ops.thin_qr(L, r)
            R = ops.project(L, Q)
            self.cores[k] = Q.reshape(self.ranks[k], self.input_shape[k], self.output_shape[k], r)
            cur = R
        self.cores[-1] = cur.reshape(self.ranks[-2], self.input_shape[-1], self.output_shape[-1], self.ranks[-1])
        self.device = flat.device
        return self

    # ---- reconstruction (tt.py:213-247) -----------------------------------------------------------------
    def reconstruct(self) -> torch.Tensor:
        if self.contract_expr is None:
            self.contract_expr = _Contraction([c.shape for c in self.cores])
        return self.contract_expr(*self.cores)

    def to_tensor(self) -> torch.Tensor:
        return self.reconstruct()

    def to_matrix(self, shape) -> torch.Tensor:
        d = self.order
        if d == 2 and _uniform(self.input_shape) and _uniform(self.output_shape) and self.cores[0].is_cuda:
            mm, nn_ = int(self.input_shape[0]), int(self.output_shape[0])
            c0, c1 = self.cores
            G1 = c0.detach().to(torch.float32).reshape(mm * nn_, -1)
            G2 = c1.detach().to(torch.float32).reshape(-1, mm * nn_)
            return ops.reconstruct2(G1, G2, min(int(shape[0]), mm * mm), min(int(shape[1]), nn_ * nn_), mm, nn_)
        if _uniform(self.input_shape) and _uniform(self.output_shape) and self.cores[0].is_cuda:
            # fused de-interleave + unpad: only the (M, N) window is written
            mm, nn_ = int(self.input_shape[0]), int(self.output_shape[0])
            out = ops.reconstruct_nd(self.cores, min(int(shape[0]), mm ** d), min(int(shape[1]), nn_ ** d), mm, nn_)
            if out is not None:
                return out
            c0 = self.cores[0]
            res = c0.detach().to(torch.float32).reshape(-1, c0.shape[-1])
            for c in self.cores[1:]:
                res = ops.matmul_rk(res, c.detach().to(torch.float32).reshape(c.shape[0], -1)).reshape(-1, c.shape[-1])
            mm, nn_ = int(self.input_shape[0]), int(self.output_shape[0])
            M = min(int(shape[0]), mm ** d)
            N = min(int(shape[1]), nn_ ** d)
            return ops.deinterleave(res.reshape(-1), M, N, mm, nn_, d, torch.float32)
        t = self.to_tensor()
        rows = 1
        for s in self.input_shape:
            rows *= int(s)
        return t.reshape(rows, -1)[: shape[0], : shape[1]]

    # ---- matricisations ------------------------------------------------------------------------------------
    def left_matrix(self, index):
        return self.cores[index].reshape((self.ranks[index] * self.input_shape[index] * self.output_shape[index], -1))

    def right_matrix(self, index):
        return self.cores[index].reshape((-1, self.input_shape[index] * self.output_shape[index] * self.ranks[index + 1]))

    def to_core(self, matrix, index):
        return matrix.reshape((self.ranks[index], self.input_shape[index], self.output_shape[index], self.ranks[index + 1]))

    # ---- compat algebra (plain PyTorch; not on the hot path) ----------------------------------------------
    def orthogonalize(self, mode="left", new_ranks=None, inplace=False):
        """tt.py:142-180."""
        if not inplace:
            tt = self.clone()
            return tt.orthogonalize(mode, new_ranks, inplace=True)
        if mode == "left":
            for k in range(self.order - 1):
                Q, S = torch.linalg.qr(self.left_matrix(k))
                W = S @ self.right_matrix(k + 1)
                if new_ranks:
                    Q = Q[:, : new_ranks[k]]
                    W = W[: new_ranks[k], :]
                self.ranks[k + 1] = Q.shape[1]
                self.cores[k] = self.to_core(Q, k)
                self.cores[k + 1] = self.to_core(W, k + 1)
        elif mode == "right":
            for k in range(self.order - 1, 0, -1):
                L = self.left_matrix(k - 1)
                Q, S = torch.linalg.qr(self.right_matrix(k).t())
                W = L @ S.t()
                if new_ranks:
                    Q = Q[:, : new_ranks[k]]
                    W = W[: new_ranks[k], :]
                self.ranks[k] = W.shape[1]
                self.cores[k - 1] = self.to_core(W, k - 1)
                self.cores[k] = self.to_core(Q.t(), k)
        return self

    def round(self, new_ranks=None, inplace=False, like=None):
        """tt.py:182-211: right-orthogonalise, then a left sweep of truncated complete QRs."""
        if isinstance(new_ranks, int):
            new_ranks = [1] + [new_ranks] * (self.order - 1) + [1]
        elif not new_ranks and not like:
            new_ranks = [1] + [i * o for i, o in zip(self.input_shape, self.output_shape)] + [1]
        elif like:
            new_ranks = like.ranks
        if not inplace:
            tt = self.clone()
            return tt.round(new_ranks, inplace=True)
        self.orthogonalize(mode="right", inplace=True)
        for k in range(self.order - 1):
            Q, S = torch.linalg.qr(self.left_matrix(k), mode="complete")
            Q = Q[:, : new_ranks[k + 1]]
            S = S[: new_ranks[k + 1], :]
            W = S @ self.right_matrix(k + 1)
            self.ranks[k] = new_ranks[k]
            self.ranks[k + 1] = new_ranks[k + 1]
            self.cores[k] = self.to_core(Q, k)
            self.cores[k + 1] = self.to_core(W, k + 1)
        return self

    def norm(self, mode="full"):
        """Squared Frobenius norm through the TT inner product (tt.py:257-260)."""
        return self.inner(self, mode=mode)

    def inner(self, other, mode="right"):
        """tt.py:262-277."""
        if mode == "full":
            env = torch.ones((1, 1), dtype=self.cores[0].dtype, device=self.cores[0].device)
            for a, b in zip(self.cores, other.cores):
                env = torch.einsum("xy,xijb,yijd->bd", env, a, b)
            return float(env.squeeze())
        a, b = self.cores[-1], other.cores[-1]
        return float(torch.einsum("aijb,aijd->bd", a, b).squeeze())

    def add_(self, constant):
        """Add a constant to every entry by appending a rank-1 block (tt.py:343-379)."""
        n_inner = 1
        for r in self.ranks:
            n_inner *= int(r)
        sub = constant / n_inner
        neg = sub < 0
        sub = abs(sub) ** (1 / self.order)
        cores = []
        for i in range(self.order):
            left = self.cores[i]
            right = torch.full_like(left, (-1 if neg else 1) * sub)
            cores.append(_block_concat(left, right, i, self.order, self.ranks[i], self.ranks[i + 1]))
        return TensorTrain.from_cores(cores)

    def __add__(self, other):
        """Block-diagonal core concatenation (tt.py:382-422)."""
        cores = []
        for i in range(self.order):
            cores.append(_block_concat(self.cores[i], other.cores[i], i, self.order, self.ranks[i], other.ranks[i + 1]))
        return TensorTrain.from_cores(cores)

    def __sub__(self, other):
        return self + (-1) * other

    def __rmul__(self, constant):
        """Spread |constant|^(1/d) over the cores (tt.py:428-447)."""
        neg = constant < 0
        sub = abs(constant) ** (1 / self.order)
        return TensorTrain.from_cores([core * ((-1 if neg else 1) * sub) for core in self.cores])

    def __mul__(self, other):
        """Element-wise product: Kronecker product of aligned cores (tt.py:449-478)."""
        cores = []
        for a, b in zip(self.cores, other.cores):
            k = torch.einsum("aijb,cijd->acijbd", a, b)
            cores.append(k.reshape(a.size(0) * b.size(0), a.size(1), a.size(2), a.size(3) * b.size(3)))
        return TensorTrain.from_cores(cores)

    def reciprocal(self):
        """tt.py:480-494 (slice-wise inverse of the interior cores)."""
        cores = []
        for i, core in enumerate(self.cores):
            if i == 0 or i == self.order - 1:
                cores.append(core.clone())
            else:
                cores.append(torch.linalg.inv(core.permute(1, 2, 0, 3)).permute(2, 0, 1, 3).contiguous())
        return TensorTrain.from_cores(cores)

    def sqrtinv(self, threshold=1e-8, max_iter=4):
        """Newton iteration for the element-wise 1/sqrt (tt.py:279-310; experimental in the reference)."""
        max_value = float(max(core.abs().max() for core in self.cores))
        prod = 1
        for r in self.ranks:
            prod *= int(r)
        max_value = prod * (max_value ** (self.order // 2))
        k = floor(log(max_value) / log(4))
        c, revc = (1 / (4 ** k)), 2 ** k
        A = c * self.clone()
        max_ranks = [1] + [i * o for i, o in zip(self.input_shape, self.output_shape)] + [1]
        while max_iter > 0:
            B = -1 / 2 * (self * (A * A).round(max_ranks)).add_(-3)
            B = B.round(max_ranks)
            C = (A * B).round(max_ranks)
            if threshold:
                if abs((C - A).norm()) < threshold:
                    return revc * C
            A = C
            max_iter -= 1
        return revc * A

    def sqrt(self, threshold=1e-3, max_iter=4):
        """Coupled Newton iteration for the element-wise sqrt (tt.py:312-341; experimental in the reference)."""
        max_value = float(self.cores[-1].abs().max())
        prod = 1
        for r in self.ranks:
            prod *= int(r)
        max_value = prod * max_value
        k = floor(log(max_value) / log(4))
        A = (1 / (4 ** k)) * self.clone()
        C = A.clone().add_(-1)
        ranks = list(A.ranks)
        while max_iter > 0 and (A - C).norm() > threshold:
            B = (A - 1 / 2 * (A * C)).round(ranks)
            D = (1 / 4 * (C * C).round(ranks) * (C.add_(-3))).round(ranks)
            max_iter -= 1
            A, C = B, D
        return 2 ** k * A


def _block_concat(left, right, i, order, left_rank_in, right_rank_out):
    if i == 0:
        return torch.cat((left, right), dim=-1)
    if i == order - 1:
        return torch.cat((left, right), dim=0)
    lp = F.pad(left, (0, right.shape[-1], 0, 0))
    rp = F.pad(right, (left.shape[-1], 0, 0, 0))
    return torch.cat([lp, rp], dim=0)
