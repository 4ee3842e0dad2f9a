"""CPU: the feature exchange in front of the contrastive loss on two gloo ranks -- one packed
all-gather, rank order, this rank's first row (keds_b200/contrastive.gather_features). The loss
itself has no CPU path and is covered by the -m gpu tests."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from keds_b200.contrastive import gather_features

        B, d = 3, 8
        I = torch.full((B, d), float(rank)) + torch.arange(B).float()[:, None]
        T = -I
        I_all, T_all, row0 = gather_features(I, T)
        ok = row0 == rank * B and I_all.shape == (world * B, d)
        for r in range(world):
            want = torch.full((B, d), float(r)) + torch.arange(B).float()[:, None]
            ok = ok and torch.equal(I_all[r * B:(r + 1) * B], want) and torch.equal(T_all[r * B:(r + 1) * B], -want)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_packed_all_gather_orders_rows_by_rank():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}


def test_single_process_returns_inputs():
    from keds_b200.contrastive import gather_features

    I, T = torch.randn(4, 8), torch.randn(4, 8)
    a, b, row0 = gather_features(I, T)
    assert a is I and b is T and row0 == 0
