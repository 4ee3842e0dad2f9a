"""Row-sharded knowledge database across the GPUs of one box: one process per GPU
(torch.distributed), every rank holds rows [lo_r, hi_r) and answers the full query batch locally
with exact fp32-re-ranked scores and GLOBAL labels; one packed all-gather of [B, k] (score, label)
blocks over NVLink (NCCL) and a merge kernel give every rank the global top-k.  The result is
bit-identical for 1, 2, 4 or 8 shards (same scores, same total order).

The reference has no such path (full replica per DDP rank: src/main.py:76,82; replicas in eval:
src/eval_retrieval.py:292,295).  The exchange is new and is the only collective on the path.

Test hooks `_local_index` / `_merge` let the CPU (gloo, world_size 2) tests drive the partition /
pack / all-gather / unpack logic with stand-ins; the product path always uses GpuIndexFlat and the
keds_topk_merge_strided kernel.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _capi
from .index import GpuIndexFlat, _stream_ptr


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """rows [lo, hi) of shard `rank`: ceil(n / world) rows each, last shards may be short/empty."""
    per = -(-n // world) if world > 0 else n
    return min(n, rank * per), min(n, (rank + 1) * per)


def packed_layout(nq: int, k: int) -> Tuple[int, int, int]:
    """(bytes per rank, byte offset of the label block, bytes of the score block) of the exchange
    buffer: [nq*k float32 scores | pad to 8 | nq*k int64 labels]."""
    d_bytes = nq * k * 4
    off_i = (d_bytes + 7) // 8 * 8
    return off_i + nq * k * 8, off_i, d_bytes


class ShardedIndex:
    def __init__(self, d: int, metric: int, device: Optional[int] = None, group=None,
                 _local_index=None, _merge: Optional[Callable] = None) -> None:
        if not dist.is_initialized():
            raise RuntimeError("ShardedIndex needs an initialised torch.distributed process group")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.d, self.metric_type = int(d), int(metric)
        self._merge = _merge
        if _local_index is not None:
            self.local = _local_index
            self.device = torch.device("cpu")
        else:
            dev = torch.cuda.current_device() if device is None else int(device)
            self.local = GpuIndexFlat(d, metric, dev)
            self.device = torch.device("cuda", dev)
        self.ntotal = 0
        self.lo = self.hi = 0

    def add_local(self, x_local, lo: int, ntotal: int) -> None:
        """This rank's rows (global row ids lo .. lo+len) of an ntotal-row database."""
        self.lo, self.hi, self.ntotal = int(lo), int(lo) + int(x_local.shape[0]), int(ntotal)
        self.local.set_id_offset(self.lo)
        if x_local.shape[0]:
            self.local.add(x_local)

    def add(self, x) -> None:
        """Every rank passes the same full matrix and keeps its own row range."""
        lo, hi = shard_bounds(int(x.shape[0]), self.world, self.rank)
        self.add_local(x[lo:hi], lo, int(x.shape[0]))

    def search(self, q: torch.Tensor, k: int):
        """q: [nq, d] float32 on this rank's device, identical on all ranks. Returns (D, I)."""
        nq = int(q.shape[0])
        total, off_i, d_bytes = packed_layout(nq, k)
        if isinstance(self.local, GpuIndexFlat):
            # the local search writes straight into the exchange buffer: no repacking copies
            send = torch.empty(total, dtype=torch.uint8, device=q.device)
            self.local.search(q, k, out=(send[:d_bytes].view(torch.float32), send[off_i:].view(torch.int64)))
        else:  # CPU test stand-in
            D, I = self.local.search(q, k)
            D, I = torch.as_tensor(D), torch.as_tensor(I)
            send = torch.empty(total, dtype=torch.uint8, device=D.device)
            send[:d_bytes].view(torch.float32).copy_(D.reshape(-1))
            send[off_i:].view(torch.int64).copy_(I.reshape(-1))
        recv = torch.empty(self.world * total, dtype=torch.uint8, device=send.device)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        if self._merge is not None:  # CPU test hook
            parts = recv.view(self.world, total)
            Dp = torch.stack([parts[r, :d_bytes].view(torch.float32).view(nq, k) for r in range(self.world)])
            Ip = torch.stack([parts[r, off_i:].view(torch.int64).view(nq, k) for r in range(self.world)])
            return self._merge(Dp, Ip, k, self.metric_type)
        lib = _capi.load()
        Dg = torch.empty((nq, k), dtype=torch.float32, device=send.device)
        Ig = torch.empty((nq, k), dtype=torch.int64, device=send.device)
        base = recv.data_ptr()
        _capi.check(
            lib.keds_topk_merge_strided(base, base + off_i, total // 4, total // 8, self.world, nq, k,
                                        self.metric_type, Dg.data_ptr(), Ig.data_ptr(),
                                        _stream_ptr(self.device.index))
        )
        return Dg, Ig
