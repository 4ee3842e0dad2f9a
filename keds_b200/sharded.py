"""Row-sharded knowledge database across the GPUs of one box: one process per GPU
(torch.distributed), every rank holds rows [lo_r, hi_r) and answers the full query batch locally
with exact fp32-re-ranked scores and GLOBAL labels; the [B, k] (score, label) blocks cross NVLink --
stored by the search kernels themselves into the peers' buffers (exchange="fused") or with one packed
NCCL all-gather (exchange="nccl") -- and a merge kernel gives every rank the global top-k.  The result is
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


class _FusedExchange:
    """Symmetric (peer-mapped) receive buffers + the native exchange handle over them. The buffer
    is allocated with torch symmetric memory, so every rank holds raw pointers into every other
    rank's copy (NVLink P2P through NVSwitch); the library lays it out (flag words, two parities
    of one slot per rank) and never sees torch types."""

    def __init__(self, world: int, rank: int, capacity: int, device: torch.device, group) -> None:
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem

        lib = _capi.load()
        slot = -(-(12 * capacity) // 48) * 48
        total = (256 + 2 * world * slot + 255) // 256 * 256
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        self.hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        peers = [int(p) for p in self.hdl.buffer_ptrs]
        assert peers[rank] == self.buf.data_ptr()
        h = C.c_void_p()
        _capi.check(lib.keds_exchange_create(world, rank, device.index, (C.c_void_p * world)(*peers), total, C.byref(h)))
        self.handle, self._lib = h, lib
        self.capacity = int(lib.keds_exchange_capacity(h))
        assert self.capacity >= capacity
        self.hdl.barrier()  # every rank's buffer is zeroed before anyone pushes

    def stats(self, stream_ptr: int) -> dict:
        import ctypes as C

        avg, mx, n, err = C.c_double(0), C.c_double(0), C.c_int64(0), C.c_uint32(0)
        _capi.check(self._lib.keds_exchange_stats(self.handle, stream_ptr, C.byref(avg), C.byref(mx), C.byref(n),
                                                  C.byref(err)))
        return {"wait_us_avg": float(avg.value), "wait_us_max": float(mx.value), "merges": int(n.value)}

    def __del__(self) -> None:
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            try:
                self._lib.keds_exchange_free(h)
            except Exception:
                pass
            self.handle = None


class ShardedIndex:
    """exchange="fused" (default on GPUs): no collective call and no copy kernel in the step -- the
    search kernels store every finished result row straight into the peers' buffers over NVLink,
    the chain's last block publishes an epoch flag and the merge kernel waits on the peers' flags
    (keds_index_search_sharded; needs torch symmetric memory and peer access inside the box).
    exchange="nccl": one packed all-gather per search, then the merge kernel."""

    def __init__(self, d: int, metric: int, device: Optional[int] = None, group=None,
                 exchange: str = "fused", _local_index=None, _merge: Optional[Callable] = None) -> None:
        if not dist.is_initialized():
            raise RuntimeError("ShardedIndex needs an initialised torch.distributed process group")
        if exchange == "p2p":
            exchange = "fused"
        if exchange not in ("nccl", "fused"):
            raise ValueError("exchange must be 'nccl' or 'fused'")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.d, self.metric_type = int(d), int(metric)
        self.exchange = exchange
        self._fx: Optional[_FusedExchange] = None
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

    def search(self, q: torch.Tensor, k: int, out=None):
        """q: [nq, d] float32 on this rank's device, identical on all ranks. Returns (D, I);
        `out=(D, I)` may name preallocated contiguous device tensors to fill (fused exchange)."""
        nq = int(q.shape[0])
        total, off_i, d_bytes = packed_layout(nq, k)
        if self.exchange == "fused" and isinstance(self.local, GpuIndexFlat):
            return self._search_fused(q, nq, k, out)
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

    def _search_fused(self, q: torch.Tensor, nq: int, k: int, out=None):
        """One native call: local search with the peer stores fused into its kernels, flag
        publication by the chain's last block, merge that waits on the peers' flags. Two buffer
        parities keep a fast rank from overwriting a block a slow rank is still merging (a rank can
        only publish step t+1 after finishing its merge of step t)."""
        lib = _capi.load()
        q = self.local._check_q_tensor(q)
        if self._fx is None or self._fx.capacity < nq * k:
            # (re)allocation is collective: every rank sees the same shapes in the same order
            if self._fx is not None:
                torch.cuda.synchronize(self.device)
                dist.barrier(group=self.group)
            self._fx = _FusedExchange(self.world, self.rank, max(nq * k, 8192), self.device, self.group)
        if out is not None:
            Dg, Ig = out
            assert Dg.is_cuda and Ig.is_cuda and Dg.is_contiguous() and Ig.is_contiguous()
            assert Dg.dtype == torch.float32 and Ig.dtype == torch.int64 and Dg.numel() == nq * k == Ig.numel()
        else:
            Dg = torch.empty((nq, k), dtype=torch.float32, device=q.device)
            Ig = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        _capi.check(lib.keds_index_search_sharded(self.local._h, self._fx.handle, q.data_ptr(), nq, k, Dg.data_ptr(),
                                                  Ig.data_ptr(), _stream_ptr(self.device.index)))
        return Dg, Ig

    def exchange_stats(self) -> dict:
        """Synchronise and report how long the merge waited for the peers (rank skew); raises if a
        peer failed to deliver within the watchdog window."""
        if self._fx is None:
            return {}
        return self._fx.stats(_stream_ptr(self.device.index))

    def check_exchange(self) -> None:
        self.exchange_stats()
