"""Data-parallel gradient exchange for SoW training (the only collective on the path).

The reference wraps the model in ``torch.nn.parallel.DistributedDataParallel`` (scripts/simple_train.py:566-572):
bucketed all-reduce(avg) of every trainable gradient, overlapped with backward.  ``SoWLinear`` works under stock
DDP unchanged.  ``FlatGradSync`` is the B200-first variant used by the trainer / bench:

  * gradients live as VIEWS into a few flat buffers (small factor grads in one bucket, the dense embedding /
    lm_head grads in their own), so one NCCL all-reduce moves a whole bucket over NVLink/NVSwitch, zeroing is one
    memset per bucket, and gradient addresses never change (which lets the fused AdamW keep its pointer table and
    makes the step CUDA-graph friendly);
  * a bucket's all-reduce is launched from a post-accumulate-grad hook as soon as its last gradient is written,
    so it overlaps the rest of backward (NCCL runs on its own stream);
  * merges stay replica-local; rank consistency of the re-initialised A is by construction (broadcast from
    rank 0 in ``accumulate``), not by RNG coincidence.

Host logic only (torch.distributed): runs on NCCL/CUDA and, for tests, on gloo/CPU.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def _dist_on() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    """What DDP's constructor does: every rank starts from rank ``src``'s parameters and buffers."""
    if not _dist_on():
        return
    for t in list(module.parameters()) + list(module.buffers()):
        if t.numel():
            dist.broadcast(t.data, src=src)


class FlatGradSync:
    """Flat-bucket gradient averaging with backward overlap."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20,
                 small_first: bool = True, overlap: bool = True, direct: Optional[Iterable[torch.nn.Parameter]] = None):
        """``direct``: parameters whose backward kernels write the gradient straight into the bucket view (the SoW
        factors: layer._grad_dst).  Their ``.grad`` is None between steps, the producing autograd node returns the
        bucket view and AccumulateGrad adopts it -- no temporary gradient, no accumulate-add launch per factor."""
        params = [p for p in params if p.requires_grad]
        self._direct = {id(p) for p in (direct or [])}
        # reverse registration order ~ order in which backward produces gradients
        order = list(reversed(params))
        self.buckets: List[dict] = []
        cur, cur_bytes, cur_key = [], 0, None
        for p in order:
            key = (p.dtype, p.device)
            nbytes = p.numel() * p.element_size()
            if cur and (key != cur_key or cur_bytes + nbytes > bucket_bytes):
                self._close(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
            cur_key = key
        if cur:
            self._close(cur)
        self.world = dist.get_world_size() if _dist_on() else 1
        self.overlap = overlap
        self._hooks = []
        self._p2b = {}
        for bi, b in enumerate(self.buckets):
            for p in b["params"]:
                self._p2b[id(p)] = bi
                if hasattr(p, "register_post_accumulate_grad_hook") and (overlap or id(p) in self._direct):
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad_ready))
        self._detach_direct()

    def _close(self, plist):
        total = sum(p.numel() for p in plist)
        flat = torch.zeros(total, dtype=plist[0].dtype, device=plist[0].device)
        off = 0
        views = []
        for p in plist:
            v = flat[off:off + p.numel()].view_as(p)
            p.grad = v                                        # gradient becomes a view into the bucket
            p._sow_grad_view = v
            views.append(v)
            off += p.numel()
        self.buckets.append({"params": plist, "flat": flat, "views": views, "ready": 0, "work": None})

    def _detach_direct(self):
        for b in self.buckets:
            for p in b["params"]:
                if id(p) in self._direct:
                    p.grad = None

    # ---- backward overlap -------------------------------------------------------------------------------
    def _on_grad_ready(self, p):
        b = self.buckets[self._p2b[id(p)]]
        v = p._sow_grad_view
        if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
            # autograd bound its own tensor (it could not adopt the bucket view): move the gradient into the bucket
            v.copy_(p.grad)
            p.grad = v
        b["ready"] += 1
        if not self.overlap:
            return
        if b["ready"] == len(b["params"]) and self.world > 1 and b["work"] is None:
            b["work"] = self._launch(b["flat"])

    def _launch(self, flat):
        if flat.is_cuda:
            return dist.all_reduce(flat, op=dist.ReduceOp.AVG, async_op=True)
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)   # gloo has no AVG
        return ("sum", work)

    def synchronize(self) -> None:
        """Finish all bucket all-reduces (launching the ones whose hooks did not fire, e.g. unused params)."""
        for b in self.buckets:
            if self.world > 1:
                if b["work"] is None:
                    b["work"] = self._launch(b["flat"])
                w = b["work"]
                if isinstance(w, tuple):
                    w[1].wait()
                    b["flat"].div_(self.world)
                else:
                    w.wait()
            b["work"] = None
            b["ready"] = 0

    def zero_grad(self) -> None:
        for b in self.buckets:
            b["flat"].zero_()
            for p, v in zip(b["params"], b["views"]):
                if id(p) in self._direct:
                    p.grad = None                                          # the backward kernels write the view directly
                elif p.grad is None or p.grad.data_ptr() != v.data_ptr():
                    p.grad = v                                             # re-attach if someone set it to None

    def bytes_per_step(self) -> int:
        return sum(b["flat"].numel() * b["flat"].element_size() for b in self.buckets)

    def remove_hooks(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def assert_replicas_consistent(tensors: Iterable[torch.Tensor], what: str = "tensor", atol: float = 0.0) -> None:
    """Raise if any tensor differs across ranks (used by the multi-GPU tests on A_new / W after a merge)."""
    if not _dist_on():
        return
    for i, t in enumerate(tensors):
        if t.numel() == 0:          # empty accumulation placeholders (acc_upweight) carry no data
            continue
        ref = t.detach().clone()
        dist.broadcast(ref, src=0)
        diff = float((ref.float() - t.detach().float()).abs().max()) if t.numel() else 0.0
        flag = torch.tensor([1.0 if diff > atol else 0.0], device=t.device if t.is_cuda else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if float(flag) > 0:
            raise RuntimeError(f"{what}[{i}] differs across ranks (max abs diff on rank {dist.get_rank()}: {diff})")
