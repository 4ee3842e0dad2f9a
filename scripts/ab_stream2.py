"""Diagnostic: streamed re-rank x slice count x wave size, same process."""
from __future__ import annotations
import json, os, sys
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

def one(ix, q, k, iters, on, S=None, W=None):
    for key, v in (("KEDS_DEBUG_SLICES", S), ("KEDS_DEBUG_WAVE", W)):
        if v is None:
            os.environ.pop(key, None)
        else:
            os.environ[key] = str(v)
    ix.set_stream_rerank(bool(on))
    ix.set_profiling(1)
    ms = timeit(lambda: ix.search(q, k), iters)
    chain = ix.profile_chain()
    ix.set_profiling(0)
    st = ix.last_stats()
    return {"on": on, "S": st["slices"], "W": st["streamed"], "ms": round(ms, 4), "score": round(chain["k_score_topk"]["ms"], 4),
            "rr": round(chain["k_select_rerank"]["ms"], 4), "err": st["err_word"]}

ix = GpuIndexFlat(D, METRIC_INNER_PRODUCT, 0); ix.add(db(50_000, 1000)); q = db(4096, 1001)
for rep in range(2):
    for on, S, W in ((0, 9, None), (0, 18, None), (1, 9, 1), (1, 18, 1), (1, 18, 2), (1, 18, 4), (1, 9, 16), (1, 18, 16), (1, 36, 1)):
        print("cfg1", json.dumps(one(ix, q, 16, 50, on, S, W)), flush=True)
del ix; torch.cuda.empty_cache()
ix = GpuIndexFlat(D, METRIC_INNER_PRODUCT, 0); ix.add(db(500_000, 1000)); q = db(16384, 1001)
for rep in range(2):
    for on, S, W in ((0, 23, None), (1, 23, 4), (1, 23, 8), (1, 23, 16), (1, 23, 32), (1, 23, 64)):
        print("16384x500k", json.dumps(one(ix, q, 16, 3, on, S, W)), flush=True)
q = db(4096, 1002)
for rep in range(2):
    for on, S, W in ((0, 23, None), (1, 23, 2), (1, 23, 4), (1, 23, 8), (1, 23, 16)):
        print("4096x500k", json.dumps(one(ix, q, 16, 10, on, S, W)), flush=True)
