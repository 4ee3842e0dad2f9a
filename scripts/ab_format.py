"""Same-process A/B of the operand format (fp16 vs bf16) on the compute-bound shapes: the scoring
kernel's rate must not depend on it; the re-rank's candidate count does.
-> gpurun_out/ab_format.json"""
from __future__ import annotations

import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keds_b200.index import METRIC_INNER_PRODUCT, GpuIndexFlat  # noqa: E402

D = 768


def db(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(n, D, generator=g, device="cuda")
    return x / x.norm(dim=1, keepdim=True)


def timeit(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


res = {}
for name, n, b, k, iters in (("cfg1_4096x50k", 50_000, 4096, 16, 50), ("B4096x500k", 500_000, 4096, 16, 10),
                             ("B128x500k", 500_000, 128, 16, 200), ("imgnet_10000x50k_k200", 50_000, 10_000, 200, 5)):
    ix = GpuIndexFlat(D, METRIC_INNER_PRODUCT, 0)
    ix.add(db(n, 1000))
    q = db(b, 1001)
    runs = []
    for fmt in ("fp16", "bf16", "fp16", "bf16"):
        ix.set_operand_format(fmt)
        ix.set_profiling(1)
        ms = timeit(lambda: ix.search(q, k), iters)
        chain = ix.profile_chain()
        ix.set_profiling(0)
        runs.append({"fmt": fmt, "ms": ms, "score_ms": chain["k_score_topk"]["ms"], "rerank_ms": chain["k_select_rerank"]["ms"],
                     "slices": ix.last_stats()["slices"]})
    res[name] = runs
    print(name, json.dumps(runs), flush=True)
    del ix
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/ab_format.json", "w"), indent=1)
