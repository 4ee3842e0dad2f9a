"""Multi-GPU check of the row-sharded index (run under torchrun, one rank per GPU):
both exchanges (NCCL all-gather, NVLink peer-memory push) against an unsharded index on rank 0's
GPU, bit for bit, plus timing.  -> gpurun_out/sharded_<world>gpu.json (rank 0)

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/test_sharded_multigpu.py
"""
from __future__ import annotations

import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keds_b200.index import METRIC_INNER_PRODUCT, METRIC_L2, GpuIndexFlat  # noqa: E402
from keds_b200.sharded import ShardedIndex, shard_bounds  # noqa: E402


def unit(n, d, seed, dev):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(n, d, generator=g, device=dev)
    return x / x.norm(dim=1, keepdim=True)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}
    ok = True
    n_rows = int(os.environ.get("KEDS_ROWS_PER_RANK", "250000"))
    for metric, mname in ((METRIC_INNER_PRODUCT, "ip"), (METRIC_L2, "l2")):
        n, d, b, k = n_rows * world, 768, 128, 64
        full = unit(n, d, 4242, dev)  # same seed on every rank: identical matrix
        q = unit(b, d, 4343, dev)
        ref = GpuIndexFlat(d, metric, local)
        ref.add(full)
        Dr, Ir = ref.search(q, k)
        ref.sync()
        del ref
        for ex in ("nccl", "p2p"):
            sh = ShardedIndex(d, metric, local, exchange=ex)
            sh.add(full)
            lo, hi = shard_bounds(n, world, rank)
            assert sh.lo == lo and sh.hi == hi
            for it in range(3):  # several steps: parities, epochs
                D, I = sh.search(q, k)
            torch.cuda.synchronize()
            same = bool(torch.equal(I, Ir) and torch.equal(D, Dr))
            ok = ok and same
            # timing
            for _ in range(5):
                sh.search(q, k)
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            steps = 300
            for _ in range(steps):
                sh.search(q, k)
            e1.record()
            dist.barrier()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if ex == "p2p":
                sh.check_exchange()
            out[f"{mname}_{ex}"] = {"bit_identical_to_unsharded": same, "ms_per_search": float(t.item()),
                                    "rows_per_rank": hi - lo, "B": b, "k": k}
            del sh
            torch.cuda.empty_cache()
        del full
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["all_ranks_ok"] = bool(flag.item())
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/sharded_{world}gpu.json", "w"), indent=1)
        print(json.dumps(out, indent=1))
    dist.destroy_process_group()
    if not out["all_ranks_ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
