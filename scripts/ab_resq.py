"""Same-process A/B of the CTA-pair scoring kernel with four resident query k-blocks (KEDS_RESQ)
against the streamed-query layout.  -> gpurun_out/ab_resq.json"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keds_b200.index import METRIC_INNER_PRODUCT, METRIC_L2, GpuIndexFlat
D = 768
def db(n, seed, d=D):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(n, d, generator=g, device="cuda")
    return x / x.norm(dim=1, keepdim=True)
res = {}
shapes = (("cfg1_4096x50k", 50_000, 4096, 16, 50, D), ("B4096x500k", 500_000, 4096, 16, 10, D), ("B16384x500k", 500_000, 16384, 16, 3, D),
          ("imgnet_k200", 50_000, 10_000, 200, 5, D), ("cfg5_4096x1M_k64", 1_000_000, 4096, 64, 5, D), ("B1000x30k_d200_l2", 30_000, 1000, 16, 20, 200))
for name, n, b, k, iters, d in shapes:
    ix = GpuIndexFlat(d, METRIC_L2 if "l2" in name else METRIC_INNER_PRODUCT, 0); ix.add(db(n, 1000, d)); q = db(b, 1001, d)
    ref = None; runs = []
    for rep in range(2):
        for v in (0, 1):
            os.environ["KEDS_RESQ_DYN"] = str(v)
            fn = lambda: ix.search(q, k)
            for _ in range(3): fn()
            torch.cuda.synchronize(); ix.set_profiling(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters): D_, I_ = fn()
            e1.record(); torch.cuda.synchronize()
            ch = ix.profile_chain(); ix.set_profiling(0); ix.sync(); st = ix.last_stats()
            if ref is None: ref = (D_.clone(), I_.clone())
            runs.append({"resq": v, "ms": round(e0.elapsed_time(e1) / iters, 4), "score_ms": round(ch["k_score_topk"]["ms"], 4),
                         "rerank_ms": round(ch["k_select_rerank"]["ms"], 4), "err": st["err_word"], "flagged": st["n_flagged"][0],
                         "same": bool(torch.equal(ref[0], D_) and torch.equal(ref[1], I_))})
            print(name, json.dumps(runs[-1]), flush=True)
    res[name] = runs
    del ix; torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/ab_resq.json", "w"), indent=1)
