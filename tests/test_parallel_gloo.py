"""CPU, world_size = 2 over gloo: the data-parallel host logic of the path.

What the N > 1 path does besides launching kernels is host logic: flat gradient buckets with backward-overlapped
all-reduce(avg) (the reference's DDP, scripts/simple_train.py:566-572), parameter broadcast at start-up, and the
replica-consistency contract of the re-initialised factors after a merge.  These tests run it with two processes.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, fn_name, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        globals()[fn_name](rank, world)
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover - reported to the parent
        import traceback
        ret[rank] = "".join(traceback.format_exception(type(e), e, e.__traceback__))
    finally:
        dist.destroy_process_group()


def _run(fn_name, world=2):
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn_name, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret.get(r) == "ok", f"rank {r}: {ret.get(r)}"


def _model(seed):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Linear(16, 32), nn.Tanh(), nn.Linear(32, 8), nn.Tanh(), nn.Linear(8, 4))


# ---- worker bodies (module level: spawn pickles them by name) ------------------------------------------------

def _body_grad_sync_matches_ddp(rank, world):
    from sow_b200.parallel import FlatGradSync, assert_replicas_consistent, broadcast_parameters
    model = _model(seed=100 + rank)                 # ranks start from DIFFERENT weights ...
    broadcast_parameters(model)                     # ... and are made identical like DDP's constructor does
    assert_replicas_consistent(model.parameters(), "param")
    ref = _model(seed=100)                          # rank 0's weights
    for p, q in zip(model.parameters(), ref.parameters()):
        assert torch.equal(p, q)
    params = list(model.parameters())
    sync = FlatGradSync(params, bucket_bytes=1 << 11, overlap=True)       # small buckets -> several of them
    assert len(sync.buckets) > 1
    opt = torch.optim.SGD(params, lr=0.1)
    for step in range(3):
        g = torch.Generator().manual_seed(7 * step + rank)                  # per-rank batches
        x = torch.randn(5, 16, generator=g)
        model(x).square().mean().backward()                                 # hooks launch the bucket all-reduces
        sync.synchronize()
        # oracle: average of the per-rank gradients, computed from scratch on every rank
        expect = [torch.zeros_like(p) for p in params]
        for r in range(world):
            m2 = _model(seed=0)
            m2.load_state_dict(model.state_dict())
            gr = torch.Generator().manual_seed(7 * step + r)
            m2(torch.randn(5, 16, generator=gr)).square().mean().backward()
            for e, q in zip(expect, m2.parameters()):
                e += q.grad / world
        for p, e in zip(params, expect):
            assert torch.allclose(p.grad, e, rtol=1e-5, atol=1e-7), (step, float((p.grad - e).abs().max()))
            assert p.grad.data_ptr() >= sync.buckets[sync._p2b[id(p)]]["flat"].data_ptr()    # still a bucket view
        opt.step()
        sync.zero_grad()
        assert all(float(p.grad.abs().max()) == 0.0 for p in params)
        assert_replicas_consistent(params, "param after step")


def _body_unused_parameter_and_no_overlap(rank, world):
    from sow_b200.parallel import FlatGradSync, assert_replicas_consistent
    torch.manual_seed(3)
    used, unused = nn.Linear(4, 4), nn.Linear(4, 4)
    params = list(used.parameters()) + list(unused.parameters())
    sync = FlatGradSync(params, overlap=False)
    x = torch.full((2, 4), float(rank + 1))
    used(x).sum().backward()
    sync.synchronize()                               # must not hang on the bucket whose hook never fired
    w_grad = used.weight.grad
    assert torch.allclose(w_grad, torch.full_like(w_grad, 2 * (1 + 2) / 2.0))        # mean over ranks of 2*(rank+1)
    assert float(unused.weight.grad.abs().max()) == 0.0
    assert_replicas_consistent([p.grad for p in params], "grad")


def _body_inconsistent_replicas_are_detected(rank, world):
    from sow_b200.parallel import assert_replicas_consistent
    t = torch.full((3,), float(rank))
    try:
        assert_replicas_consistent([t], "A_new")
    except RuntimeError as e:
        assert "differs across ranks" in str(e)
    else:
        raise AssertionError("rank-dependent tensor was not flagged")
    # the fix the merge applies (sow_b200/layer.py:_reinit): broadcast from rank 0
    dist.broadcast(t, src=0)
    assert_replicas_consistent([t], "A_new")


# ---- tests ------------------------------------------------------------------------------------------------------

def test_flat_grad_sync_equals_ddp_average_world2():
    _run("_body_grad_sync_matches_ddp")


def test_unused_parameters_and_non_overlapped_mode_world2():
    _run("_body_unused_parameter_and_no_overlap")


def test_rank_divergent_reinit_is_detected_and_fixed_by_broadcast_world2():
    _run("_body_inconsistent_replicas_are_detected")
