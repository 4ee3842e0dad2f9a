"""Same-process A/B of the streamed re-rank (re-rank blocks start on their query tile's counter,
under the scoring of later tiles) against the plain chain, on the compute-bound shapes.
-> gpurun_out/ab_stream.json"""
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


want = set(sys.argv[1:])
res = {}
for name, n, b, k, iters in (("cfg1_4096x50k", 50_000, 4096, 16, 50), ("B4096x500k", 500_000, 4096, 16, 10),
                             ("cfg3_65536x500k", 500_000, 65536, 16, 2), ("imgnet_10000x50k_k200", 50_000, 10_000, 200, 5),
                             ("cfg5_4096x1M_k64", 1_000_000, 4096, 64, 5), ("B1024x500k", 500_000, 1024, 16, 20)):
    if want and name not in want:
        continue
    ix = GpuIndexFlat(D, METRIC_INNER_PRODUCT, 0)
    ix.add(db(n, 1000))
    q = db(b, 1001)
    runs = []
    ref = None
    for on in (1, 0, 1, 0):
        ix.set_stream_rerank(bool(on))
        ix.set_profiling(1)
        ms = timeit(lambda: ix.search(q, k), iters)
        chain = ix.profile_chain()
        ix.set_profiling(0)
        Dq, Iq = ix.search(q, k)
        ix.sync()
        st = ix.last_stats()
        if ref is None:
            ref = (Dq.clone(), Iq.clone())
        same = bool(torch.equal(ref[0], Dq) and torch.equal(ref[1], Iq))
        runs.append({"streamed": st["streamed"], "ms": round(ms, 4), "score_ms": round(chain["k_score_topk"]["ms"], 4),
                     "rerank_ms": round(chain["k_select_rerank"]["ms"], 4), "slices": st["slices"],
                     "flagged": st["n_flagged"][0], "err": st["err_word"], "same_as_first": same})
    res[name] = runs
    print(name, json.dumps(runs), flush=True)
    del ix
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/ab_stream.json", "w"), indent=1)
