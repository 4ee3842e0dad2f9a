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


class _PeerExchange:
    """Symmetric (peer-mapped) exchange buffers for one (nq, k) shape: two parities of
    [world] result blocks + per-parity flag words, allocated with torch symmetric memory so that
    every rank holds raw pointers into every other rank's copy (NVLink P2P)."""

    def __init__(self, world: int, rank: int, slot_bytes: int, device: torch.device, group) -> None:
        import torch.distributed._symmetric_memory as symm_mem

        self.world, self.rank, self.slot = world, rank, slot_bytes
        self.flags_off = 2 * world * slot_bytes
        total = self.flags_off + 2 * world * 4
        total = (total + 255) // 256 * 256
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        self.hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.peer_base = [int(p) for p in self.hdl.buffer_ptrs]
        assert self.peer_base[rank] == self.buf.data_ptr()
        self.ticket = torch.zeros(1, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.epoch = 0
        self.ptrs = [None, None]  # per parity: (ctypes peer block pointers, ctypes peer flag pointers)
        self.hdl.barrier()  # every rank's buffer is zeroed before anyone pushes

    def region(self, parity: int, r: int) -> int:
        return (parity * self.world + r) * self.slot


class ShardedIndex:
    """exchange="nccl": one packed all-gather per search (default).
    exchange="p2p":  no collective call in the step -- every rank stores its block straight into
    the peers' buffers over NVLink and the merge kernel waits on epoch flags (needs torch symmetric
    memory and peer access between the GPUs of the box)."""

    def __init__(self, d: int, metric: int, device: Optional[int] = None, group=None,
                 exchange: str = "nccl", _local_index=None, _merge: Optional[Callable] = None) -> None:
        if not dist.is_initialized():
            raise RuntimeError("ShardedIndex needs an initialised torch.distributed process group")
        if exchange not in ("nccl", "p2p"):
            raise ValueError("exchange must be 'nccl' or 'p2p'")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.d, self.metric_type = int(d), int(metric)
        self.exchange = exchange
        self._px = {}  # (nq, k) -> _PeerExchange
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
        if self.exchange == "p2p" and isinstance(self.local, GpuIndexFlat) and self.world > 1:
            return self._search_p2p(q, nq, k, total, off_i, d_bytes)
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

    def _search_p2p(self, q: torch.Tensor, nq: int, k: int, total: int, off_i: int, d_bytes: int):
        """local search -> k_p2p_push (stores into every peer + epoch flag) -> merge that waits on
        the peers' flags. Parities alternate so a fast rank never overwrites a block a slow rank is
        still merging (a rank can only publish step t+1 after finishing its merge of step t)."""
        import ctypes as C

        lib = _capi.load()
        slot = (total + 15) // 16 * 16
        px = self._px.get((nq, k))
        if px is None:
            px = self._px[(nq, k)] = _PeerExchange(self.world, self.rank, slot, q.device, self.group)
        px.epoch += 1
        parity = px.epoch & 1
        mine = px.region(parity, self.rank)
        blk = px.buf[mine:mine + total]
        self.local.search(q, k, out=(blk[:d_bytes].view(torch.float32), blk[off_i:].view(torch.int64)))
        stream = _stream_ptr(self.device.index)
        if px.ptrs[parity] is None:  # pointer tables are fixed per parity: build them once
            vp = C.c_void_p * self.world
            px.ptrs[parity] = (
                vp(*[px.peer_base[r] + mine for r in range(self.world)]),
                vp(*[px.peer_base[r] + px.flags_off + (parity * self.world + self.rank) * 4
                     for r in range(self.world)]),
            )
        dst, flg = px.ptrs[parity]
        _capi.check(lib.keds_p2p_push(px.buf.data_ptr() + mine, slot, dst, flg, self.world, self.rank, px.epoch,
                                      px.ticket.data_ptr(), stream))
        Dg = torch.empty((nq, k), dtype=torch.float32, device=q.device)
        Ig = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        base = px.buf.data_ptr() + px.region(parity, 0)
        flags = px.buf.data_ptr() + px.flags_off + parity * self.world * 4
        _capi.check(
            lib.keds_topk_merge_wait(base, base + off_i, slot // 4, slot // 8, self.world, nq, k, self.metric_type,
                                     Dg.data_ptr(), Ig.data_ptr(), flags, self.rank, px.epoch, px.err.data_ptr(),
                                     stream)
        )
        return Dg, Ig

    def check_exchange(self) -> None:
        """Raise if a peer failed to deliver within the watchdog window of any p2p merge so far."""
        for px in self._px.values():
            e = int(px.err.item())
            if e:
                raise RuntimeError(f"p2p exchange: peer {e - 0x500} did not deliver (error word {e:#x})")
