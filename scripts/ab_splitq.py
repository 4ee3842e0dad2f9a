"""Same-process A/B of the single-CTA scoring kernel's staging layout on config 2 (128 queries,
2 x 0.5M rows, retrieve2): four combined 48-KB stages vs five 32-KB row stages + a query ring of
its own (KEDS_SPLIT_Q).  -> gpurun_out/ab_splitq.json"""
from __future__ import annotations

import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keds_b200 import retrieval as kr  # noqa: E402
from keds_b200.index import METRIC_INNER_PRODUCT, GpuIndexFlat  # noqa: E402

D = 768


def db(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(n, D, generator=g, device="cuda")
    return x / x.norm(dim=1, keepdim=True)


rows = [db(500_000, 1002), db(500_000, 1003)]
q = db(128, 1004)
pairs = {}
for mode in ("0", "1"):
    os.environ["KEDS_SPLIT_Q"] = mode
    ia, ib = GpuIndexFlat(D, METRIC_INNER_PRODUCT, 0), GpuIndexFlat(D, METRIC_INNER_PRODUCT, 0)
    ia.add(rows[0])
    ib.add(rows[1])
    bufs = {}
    kr.retrieve2(ia, ib, q, 16, want_feats=True, pool_mode=kr.POOL_SOFTMAX, out=bufs)   # reads the env
    ia.sync()
    pairs[mode] = (ia, ib, bufs)
ref = pairs["0"][2]
same = torch.equal(ref["I_img"], pairs["1"][2]["I_img"]) and torch.equal(ref["D_txt"], pairs["1"][2]["D_txt"])
res = {"identical_results": bool(same), "runs": []}
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
for rep in range(3):
    for mode in ("0", "1"):
        ia, ib, bufs = pairs[mode]
        fn = lambda: kr.retrieve2(ia, ib, q, 16, want_feats=True, pool_mode=kr.POOL_SOFTMAX, out=bufs)
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        ia.set_profiling(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        chain = ia.profile_chain()
        ia.set_profiling(0)
        r = {"split_q": int(mode), "us_per_step": e0.elapsed_time(e1) / steps * 1e3,
             "score_us": chain["k_score_topk"]["ms"] * 1e3, "rerank_us": chain["k_select_rerank"]["ms"] * 1e3}
        res["runs"].append(r)
        print(r, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/ab_splitq.json", "w"), indent=1)
print("identical_results", same)
