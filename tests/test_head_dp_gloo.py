"""world_size-2 `gloo` test of the training step's gradient all-reduce (BASELINE config 3's only collective):
two ranks with identical head replicas and half the batch each must end up with exactly the gradient — and, after
an AdamW step, the parameters — of one process seeing the whole batch."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "vla-from-fastvlm_b200"))


def _head(seed=0):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Linear(24, 32), nn.LayerNorm(32), nn.SiLU(), nn.Linear(32, 5))


def _data():
    g = torch.Generator().manual_seed(7)
    return torch.randn(8, 24, generator=g), torch.randn(8, 5, generator=g)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from vla_fastvlm.training.head_dp import HeadGradAllReduce

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        head = _head()
        red = HeadGradAllReduce(head.parameters())
        opt = torch.optim.AdamW(head.parameters(), lr=1e-2)
        x, y = _data()
        n = x.shape[0] // world
        xs, ys = x[rank * n:(rank + 1) * n], y[rank * n:(rank + 1) * n]
        red.zero()
        nn.functional.mse_loss(head(xs), ys).backward()
        red.all_reduce()
        grads = red.flat.clone()
        opt.step()
        q.put((rank, grads.tolist(), torch.cat([p.detach().flatten() for p in head.parameters()]).tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_head_gradients_match_single_process():
    from vla_fastvlm.training.head_dp import HeadGradAllReduce

    # single-process reference on the whole batch
    head = _head()
    red = HeadGradAllReduce(head.parameters())
    assert red.numel == sum(p.numel() for p in head.parameters())
    assert all(p.grad.data_ptr() >= red.flat.data_ptr() for p in head.parameters())
    opt = torch.optim.AdamW(head.parameters(), lr=1e-2)
    x, y = _data()
    red.zero()
    nn.functional.mse_loss(head(x), y).backward()
    red.all_reduce()  # world size 1: no collective
    ref_g = red.flat.clone()
    opt.step()
    ref_p = torch.cat([p.detach().flatten() for p in head.parameters()])

    world, port = 2, 29577
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, g, pv in res:
        assert torch.allclose(torch.tensor(g), ref_g, atol=1e-6, rtol=1e-5)
        assert torch.allclose(torch.tensor(pv), ref_p, atol=1e-6, rtol=1e-5)
    assert res[0][1] == res[1][1]  # every rank holds the same averaged gradient


def test_reducer_detects_detached_gradients():
    import pytest
    from vla_fastvlm.training.head_dp import HeadGradAllReduce

    head = _head()
    red = HeadGradAllReduce(head.parameters())
    torch.optim.SGD(head.parameters(), lr=0.1).zero_grad(set_to_none=True)
    x, y = _data()
    nn.functional.mse_loss(head(x), y).backward()
    with pytest.raises(RuntimeError, match="flat buffer"):
        red.all_reduce()
    g = red.clip_(1e-3)
    assert red.grad_norm() <= 1e-3 + 1e-9 or float(g) == 0.0
