"""CPU, world_size 2, gloo: the row-shard / pack / all-gather / unpack / merge plumbing of
ShardedIndex with stand-in local search (the oracle) and a numpy merge. The CUDA local search and
merge kernel are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import knn_oracle as orc


class OracleLocal:
    def __init__(self, metric):
        self.metric = "l2" if metric == 1 else "ip"
        self.base = np.zeros((0, 0), np.float32)
        self.off = 0

    def set_id_offset(self, off):
        self.off = off

    def add(self, x):
        self.base = np.asarray(x, np.float32)

    def search(self, q, k):
        if self.base.shape[0] == 0:
            return orc.search(np.zeros((0, q.shape[1]), np.float32), q.numpy(), k, self.metric)
        D, I = orc.search(self.base, q.numpy(), k, self.metric)
        return D, np.where(I >= 0, I + self.off, I)


def numpy_merge(Dp, Ip, k, metric):
    Dp, Ip = Dp.numpy(), Ip.numpy()
    R, nq, _ = Dp.shape
    D = np.full((nq, k), orc.FLT_MAX if metric == 1 else -orc.FLT_MAX, np.float32)
    I = np.full((nq, k), -1, np.int64)
    for q in range(nq):
        d, i = Dp[:, q].reshape(-1), Ip[:, q].reshape(-1)
        keep = i >= 0
        d, i = d[keep], i[keep]
        order = np.lexsort((i, d if metric == 1 else -d))[:k]
        D[q, :len(order)] = d[order]
        I[q, :len(order)] = i[order]
    return D, I


def _worker(rank, world, port, n, metric, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from keds_b200.sharded import ShardedIndex
        rng = np.random.default_rng(5)
        db = rng.standard_normal((n, 32)).astype(np.float32)
        q = torch.from_numpy(rng.standard_normal((9, 32)).astype(np.float32))
        ix = ShardedIndex(32, metric, _local_index=OracleLocal(metric), _merge=numpy_merge)
        ix.add(db)
        D, I = ix.search(q, 6)
        Dr, Ir = orc.search(db, q.numpy(), 6, "l2" if metric == 1 else "ip")
        ok = np.array_equal(I, Ir) and np.allclose(D, Dr, atol=1e-6)
        out[rank] = bool(ok) and ix.ntotal == n
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(n, metric):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), n, metric, out), nprocs=2, join=True)
    assert out.get(0) is True and out.get(1) is True, dict(out)


def test_two_shards_equal_one_index_ip():
    _run(101, 0)


def test_two_shards_equal_one_index_l2():
    _run(64, 1)


def test_fewer_rows_than_k_per_shard():
    _run(5, 0)  # shards of 3 and 2 rows, k = 6 > ntotal: -1 padding must survive the merge
