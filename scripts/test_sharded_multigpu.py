"""Multi-GPU check of the row-sharded index (run under torchrun, one rank per GPU): the fused
NVLink exchange (search kernels store result rows into the peers, flag, merge-wait) and the NCCL
all-gather variant against an unsharded index on every rank's GPU, bit for bit, plus timing.
Covers: both metrics, k = 16 / 64, one and several query tiles, a batch above one pass (17,000),
queries that fail the certificate (answered and pushed by the exact fallback), exact-only shards
(fewer rows than the planner wants) and an empty shard.  -> gpurun_out/sharded_<world>gpu.json

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
    steps = int(os.environ.get("KEDS_STEPS", "300"))

    def compare(tag, sh, ref, q, k, reps=3):
        nonlocal ok
        Dr, Ir = ref.search(q, k)
        ref.sync()
        for _ in range(reps):  # several steps: parities, epochs
            D, I = sh.search(q, k)
        torch.cuda.synchronize()
        sh.local.sync()  # the status words of the last local search reach last_stats()
        same = bool(torch.equal(I, Ir) and torch.equal(D, Dr))
        ok = ok and same
        out[tag] = {"bit_identical_to_unsharded": same, "B": int(q.shape[0]), "k": k,
                    "local_flagged": sh.local.last_stats()["n_flagged"][0], "exact_only": sh.local.last_stats()["exact_only"]}
        return same

    for metric, mname in ((METRIC_INNER_PRODUCT, "ip"), (METRIC_L2, "l2")):
        n, d = n_rows * world, 768
        full = unit(n, d, 4242, dev)  # same seed on every rank: identical matrix
        ref = GpuIndexFlat(d, metric, local)
        ref.add(full)
        for ex in ("fused", "nccl"):
            sh = ShardedIndex(d, metric, local, exchange=ex)
            sh.add(full)
            lo, hi = shard_bounds(n, world, rank)
            assert sh.lo == lo and sh.hi == hi
            q = unit(128, d, 4343, dev)
            compare(f"{mname}_{ex}_B128_k64", sh, ref, q, 64)
            compare(f"{mname}_{ex}_B300_k16", sh, ref, unit(300, d, 4344, dev), 16)
            if metric == METRIC_INNER_PRODUCT:
                compare(f"{mname}_{ex}_B17000_k16", sh, ref, unit(17000, d, 4345, dev), 16, reps=2)
                # queries whose certificate fails locally: the exact fallback answers and pushes them
                sh.local.set_eps_scale(300.0)
                compare(f"{mname}_{ex}_B128_k16_all_flagged", sh, ref, q, 16, reps=2)
                assert sh.local.last_stats()["n_flagged"][0] == 128
                sh.local.set_eps_scale(1.0)
            # timing at the configs[4] shape
            for _ in range(5):
                sh.search(q, 64)
            if ex == "fused":
                sh.exchange_stats()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                sh.search(q, 64)
            e1.record()
            dist.barrier()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ent = {"ms_per_search": float(t.item()), "rows_per_rank": hi - lo, "B": 128, "k": 64}
            if ex == "fused":
                ent["merge_wait_for_peers_us"] = sh.exchange_stats()
            out[f"{mname}_{ex}_timing"] = ent
            # the local search alone, same shape, for the exchange's share
            lt = []
            for _ in range(2):
                torch.cuda.synchronize()
                l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                l0.record()
                for _ in range(steps):
                    sh.local.search(q, 64)
                l1.record()
                torch.cuda.synchronize()
                lt.append(l0.elapsed_time(l1) / steps)
            out[f"{mname}_{ex}_timing"]["local_search_alone_ms"] = min(lt)
            del sh
            torch.cuda.empty_cache()
        del full, ref
        torch.cuda.empty_cache()

    # tiny database: 3 * world + 1 rows -> short shards (exact-only planner branch) and, with
    # world > 4 rows missing, an EMPTY last shard
    for n_small in (3 * world + 1, max(1, world - 1)):
        small = unit(n_small, 64, 777, dev)
        ref = GpuIndexFlat(64, METRIC_INNER_PRODUCT, local)
        ref.add(small)
        sh = ShardedIndex(64, METRIC_INNER_PRODUCT, local, exchange="fused")
        sh.add(small)
        compare(f"tiny_{n_small}_rows_fused", sh, ref, unit(9, 64, 778, dev), 5)
        del sh, ref

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
