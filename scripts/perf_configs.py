"""Timing of every BASELINE.json config shape on one GPU (device-resident queries), with the
scoring kernel's own duration (in-kernel timer) beside the whole call, and a same-process A/B of
programmatic dependent launch.  -> gpurun_out/perf_configs.json

    python scripts/perf_configs.py [cfg2 cfg1 cfg3 cfg5 ...]
"""
from __future__ import annotations

import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keds_b200 import retrieval as kr  # noqa: E402
from keds_b200.index import METRIC_INNER_PRODUCT, GpuIndexFlat, search2  # noqa: E402

HBM = 6544e9
TENSOR = 1649.1e12
D = 768


def db(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(n, D, generator=g, device="cuda")
    return x / x.norm(dim=1, keepdim=True)


def timeit(fn, iters, warm=5):
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


def roof_ms(b, n, k, ndb):
    flops = 2.0 * b * n * D * ndb
    byts = (n * D * 2 + 12 * b * k) * ndb + b * D * 2
    return max(flops / TENSOR, byts / HBM) * 1e3, ("tensor" if flops / TENSOR > byts / HBM else "hbm")


def run(name, ix_list, q, k, iters, fn):
    out = {}
    a = ix_list[0]
    for pdl in (1, 0, 1, 0):
        a.set_pdl(bool(pdl))
        a.set_profiling(1)
        ms = timeit(fn, iters)
        chain = a.profile_chain()
        a.set_profiling(0)
        key = f"pdl{pdl}"
        out.setdefault(key, []).append({"ms": ms, "score_kernel_ms": chain["k_score_topk"]["ms"], "chain": chain})
    a.set_pdl(True)
    b = q.shape[0]
    n = a.ntotal
    r_ms, bound = roof_ms(b, n, k, len(ix_list))
    best = min(x["ms"] for v in out.values() for x in v)
    res = {"shape": {"B": b, "N": n, "k": k, "dbs": len(ix_list)}, "roofline_ms": r_ms, "bound": bound,
           "best_ms": best, "frac_of_roofline": r_ms / best, "qps": b / best * 1e3, "runs": out, "stats": a.last_stats()}
    print(name, json.dumps(res), flush=True)
    return res


def cfg2():
    ia, ib = GpuIndexFlat(D, METRIC_INNER_PRODUCT, 0), GpuIndexFlat(D, METRIC_INNER_PRODUCT, 0)
    ia.add(db(500_000, 1002))
    ib.add(db(500_000, 1003))
    q = db(128, 1004)
    bufs = {}
    r = {}
    r["cfg2_search2"] = run("cfg2_search2", [ia, ib], q, 16, 300, lambda: search2(ia, ib, q, 16))
    r["cfg2_retrieve2"] = run("cfg2_retrieve2", [ia, ib], q, 16, 300,
                              lambda: kr.retrieve2(ia, ib, q, 16, want_feats=True, pool_mode=kr.POOL_SOFTMAX, out=bufs))
    r["cfg2_single_db"] = run("cfg2_single_db", [ia], q, 16, 300, lambda: ia.search(q, 16))
    q3 = db(16384, 1005)
    r["cfg3_16k_of_65536"] = run("cfg3_16k", [ia], q3, 16, 5, lambda: ia.search(q3, 16))
    q4 = db(4096, 1006)
    r["B4096_500k"] = run("B4096_500k", [ia], q4, 16, 10, lambda: ia.search(q4, 16))
    q5 = db(1024, 1007)
    r["B1024_500k"] = run("B1024_500k", [ia], q5, 16, 20, lambda: ia.search(q5, 16))
    return r


def cfg1():
    ia = GpuIndexFlat(D, METRIC_INNER_PRODUCT, 0)
    ia.add(db(50_000, 1000))
    q = db(4096, 1001)
    r = {"cfg1": run("cfg1", [ia], q, 16, 50, lambda: ia.search(q, 16))}
    g = GpuIndexFlat(D, METRIC_INNER_PRODUCT, 0)
    g.add(db(50_000, 1008))
    q2 = db(10_000, 1009)
    r["cfg4_imgnet_k200"] = run("cfg4_imgnet_k200", [g], q2, 200, 10, lambda: g.search(q2, 200))
    return r


def cfg5():
    ia = GpuIndexFlat(D, METRIC_INNER_PRODUCT, 0)
    ia.add(db(1_000_000, 1010))
    q = db(128, 1020)
    r = {"cfg5_shard_1M_k64_B128": run("cfg5_1M_k64", [ia], q, 64, 100, lambda: ia.search(q, 64))}
    q2 = db(4096, 1021)
    r["cfg5_shard_1M_k64_B4096"] = run("cfg5_1M_k64_B4096", [ia], q2, 64, 5, lambda: ia.search(q2, 64))
    return r


if __name__ == "__main__":
    want = sys.argv[1:] or ["cfg2", "cfg1", "cfg5"]
    res = {}
    for w in want:
        res.update({"cfg2": cfg2, "cfg1": cfg1, "cfg5": cfg5}[w]())
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/perf_configs.json", "w"), indent=1)
