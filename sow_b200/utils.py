"""Math helpers -- host-side mirror of ``tn_gradient.utils`` (reference file tn_gradient/utils.py).

``qr_weight`` / ``pad_matrix`` / ``unpad_matrix`` are on the hot path and run on the CUDA kernels when given CUDA
tensors; the random-matrix / unfolding / printing helpers are diagnostics kept in plain PyTorch (SURVEY.md 2,
row 5: out of scope for the kernels, provided for import compatibility).
"""
from __future__ import annotations

from math import ceil

import torch

from . import ops

_KERNEL_QR_MAX_RANK = 256


def qr_weight(weight: torch.Tensor, rank: int = None):
    """Truncated QR in fp32, results cast back to ``weight.dtype`` (tn_gradient/utils.py:8-30).

    On CUDA with ``rank`` <= 256:  Q = thin-QR kernel on the first ``rank`` columns, R = Q^T W (projection kernel)
    -- the same Q[:, :rank], R[:rank, :] as the full Householder QR up to the per-column sign (here diag(R) >= 0).
    Otherwise (no rank / huge rank / CPU tensor): torch.linalg.qr, as in the reference.
    """
    orig_dtype = weight.dtype
    w = weight if weight.dtype == torch.float32 else weight.to(torch.float32)
    if rank and w.is_cuda and rank <= _KERNEL_QR_MAX_RANK and rank <= w.shape[0] and w.dim() == 2:
        w = w.contiguous()
        Q = ops.thin_qr(w, rank)
        R = ops.project(w, Q)
    else:
        Q, R = torch.linalg.qr(w)
        if rank:
            Q = Q[:, :rank]
            R = R[:rank, :]
    if orig_dtype != torch.float32:
        Q = Q.type(orig_dtype)
        R = R.type(orig_dtype)
    return Q, R


def svd_weight(weight: torch.Tensor, rank: int = None):
    """Truncated SVD in fp32 (tn_gradient/utils.py:32-57); diagnostic only (export_alignment), stays on torch."""
    orig_dtype = weight.dtype
    w = weight if weight.dtype == torch.float32 else weight.to(torch.float32)
    U, S, V = torch.linalg.svd(w)
    if rank:
        U, S, V = U[:, :rank], S[:rank], V[:rank, :]
    if orig_dtype != torch.float32:
        U, S, V = U.type(orig_dtype), S.type(orig_dtype), V.type(orig_dtype)
    return U, S, V


def pad_matrix(matrix: torch.Tensor, new_shape):
    """Zero-pad to ``new_shape`` in the default dtype (tn_gradient/utils.py:78-84)."""
    padded = torch.zeros(new_shape, device=matrix.device)
    padded[: matrix.shape[0], : matrix.shape[1]] = matrix
    return padded


def unpad_matrix(matrix: torch.Tensor, shape):
    return matrix[: shape[0], : shape[1]]


def closest_factorization(n: int, d: int):
    """Greedy d-factor split of n used by tests/tt_adam_update.py (tn_gradient/utils.py:89-99)."""
    factors = []
    prod, orig = 1, n
    while n > 1:
        k = ceil(n ** (1 / d))
        factors.append(k)
        n, prod, d = n // k, prod * k, d - 1
        if n == 1:
            if prod < orig:
                factors[-1] += n
            return factors, prod
    return factors, prod


def randhaar(n: int) -> torch.Tensor:
    """Haar-distributed orthogonal matrix (QR of a Gaussian with sign fix)."""
    q, r = torch.linalg.qr(torch.randn(n, n, dtype=torch.float64))
    return (q * torch.sign(torch.diagonal(r))).to(torch.float32)


def randuptri(n: int, scale: float = 1.0) -> torch.Tensor:
    R = torch.triu(torch.randn(n, n))
    for i in range(n):
        R[i, i] = torch.sqrt(torch.distributions.Chi2(df=n - i).sample()) * scale
    return R


def perturbe_random(matrix: torch.Tensor, scale: float = 0.02) -> torch.Tensor:
    return matrix + torch.randn(matrix.size(), device=matrix.device) * scale


def generate_rank_k(shape, rank, mix=1, pos=False):
    """Sum of ``mix`` random rank-``rank`` CP tensors."""
    tensor = torch.zeros(tuple(shape))
    letters = "abcdefghijklmnopqrstuvwxy"
    eq = ",".join(f"{letters[i]}z" for i in range(len(shape))) + "->" + letters[: len(shape)]
    for _ in range(mix):
        factors = [torch.rand(dim, rank) for dim in shape]
        if not pos:
            factors = [2 * f - 1 for f in factors]
        tensor += torch.einsum(eq, *factors)
    return tensor


def unfolding(tensor: torch.Tensor, mode: int) -> torch.Tensor:
    d = tensor.dim()
    if mode < 0:
        mode = d + mode
    if mode < 0 or mode >= d:
        raise ValueError("Mode must be between 1 - d and d + 1, d being the number of dimensions of the tensor")
    return torch.reshape(torch.moveaxis(tensor, mode, 0), (tensor.shape[mode], -1))


def left_unfolding(tensor):
    return unfolding(tensor, -1).t()


def right_unfolding(tensor):
    return unfolding(tensor, 0)


def __colorized_str__(self) -> str:
    """Module pretty-printer that scripts/simple_train.py:45-46 monkey-patches onto nn.Module.__str__: like
    nn.Module.__repr__, with the trainable / frozen parameter counts of every sub-module appended."""
    def fmt(mod, indent):
        own = list(mod.parameters(recurse=True))
        train = sum(p.numel() for p in own if p.requires_grad)
        frozen = sum(p.numel() for p in own if not p.requires_grad)
        head = f"{mod._get_name()}({mod.extra_repr()})" if not mod._modules else f"{mod._get_name()}("
        tag = f"  [trainable {train:,} | frozen {frozen:,}]"
        if not mod._modules:
            return head + tag
        lines = [head + tag]
        for name, child in mod._modules.items():
            if child is None:
                continue
            lines.append(" " * (indent + 2) + f"({name}): " + fmt(child, indent + 2))
        lines.append(" " * indent + ")")
        return "\n".join(lines)

    return fmt(self, 0)
