"""GPU bring-up: staged checks of the native path against the numpy oracle, most basic first.
Run on a B200 box:  python scripts/bringup.py [stage ...]   -> gpurun_out/bringup.json"""
from __future__ import annotations

import json
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keds_b200 import index as kx  # noqa: E402
from oracle import knn_oracle as orc  # noqa: E402

OUT = {}


def unit(n, d, seed, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g)
    return (x / x.norm(dim=1, keepdim=True)).to(device)


def stage_gemm():
    """raw tensor-core scores vs fp32 matmul of the bf16-rounded operands"""
    res = {}
    for (n, b, d) in [(5000, 200, 768), (300, 7, 64), (1000, 128, 200)]:
        db = unit(n, d, 1)
        q = unit(b, d, 2)
        ix = kx.GpuIndexFlat(d, kx.METRIC_INNER_PRODUCT, 0)
        ix.add(db.numpy())
        got = ix.debug_scores(q.cuda()).cpu()
        ref = (q.bfloat16().float().double() @ db.bfloat16().float().double().t()).float()
        err = (got - ref).abs().max().item()
        exact = (q.double() @ db.double().t()).float()
        res[f"{n}x{b}x{d}"] = {"max_err_vs_bf16_ref": err, "max_err_vs_exact": (got - exact).abs().max().item(),
                               "stats": ix.last_stats()}
        print("gemm", n, b, d, res[f"{n}x{b}x{d}"], flush=True)
    return res


def _cmp(name, ix, db, q, k, metric, flags=0):
    t0 = time.time()
    D, I = ix.search(q, k, flags)
    dt = time.time() - t0
    Dr, Ir = orc.search(db, q, k, metric)
    c = orc.compare_topk(Dr, Ir, D, I, db, q, metric)
    c["stats"] = ix.last_stats()
    c["wall_s"] = dt
    print(name, c, flush=True)
    return c


def stage_search_small():
    res = {}
    for metric, mname in [(kx.METRIC_INNER_PRODUCT, "ip"), (kx.METRIC_L2, "l2")]:
        for (n, b, d, k) in [(5000, 200, 768, 16), (70000, 130, 768, 16), (999, 5, 768, 16), (3000, 64, 96, 4)]:
            db = unit(n, d, 11).numpy()
            if mname == "l2":
                db = db * np.linspace(0.9, 1.1, n, dtype=np.float32)[:, None]
            q = unit(b, d, 12).numpy()
            ix = kx.GpuIndexFlat(d, metric, 0)
            ix.add(db)
            res[f"{mname}_{n}x{b}x{d}_k{k}"] = _cmp(f"search {mname} {n} {b} {d} {k}", ix, db, q, k, mname)
            res[f"{mname}_{n}x{b}x{d}_k{k}_exact"] = _cmp(f"exact  {mname} {n} {b} {d} {k}", ix, db, q, k, mname, 1)
    return res


def stage_cfg1():
    db = unit(50000, 768, 1000).numpy()
    q = unit(4096, 768, 1001).numpy()
    ix = kx.GpuIndexFlat(768, kx.METRIC_INNER_PRODUCT, 0)
    ix.add(db)
    return {"cfg1": _cmp("cfg1", ix, db, q, 16, "ip")}


def stage_time():
    """first timing of the training-step shape: 128 queries, two 500k x 768 databases"""
    res = {}
    n, d, b, k = 500000, 768, 128, 16
    dbs = []
    for s in (1002, 1003):
        g = torch.Generator(device="cuda").manual_seed(s)
        x = torch.randn(n, d, generator=g, device="cuda")
        dbs.append(x / x.norm(dim=1, keepdim=True))
    q = unit(b, d, 1004, "cuda")
    ia = kx.GpuIndexFlat(d, kx.METRIC_INNER_PRODUCT, 0)
    ib = kx.GpuIndexFlat(d, kx.METRIC_INNER_PRODUCT, 0)
    ia.add(dbs[0])
    ib.add(dbs[1])
    for name, fn in [("search1", lambda: ia.search(q, k)), ("search2", lambda: kx.search2(ia, ib, q, k))]:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        iters = 20
        for _ in range(iters):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        ia.sync()
        ms = ev0.elapsed_time(ev1) / iters
        res[name] = {"ms": ms, "qps": b / ms * 1e3, "stats": ia.last_stats()}
        print(name, res[name], flush=True)
    # parity on a query subset against torch fp32 on the GPU (oracle-free sanity at full size)
    (Da, Ia), (Db, Ib) = kx.search2(ia, ib, q, k)
    ia.sync()
    for nm, dbt, Dg, Ig in [("img", dbs[0], Da, Ia), ("txt", dbs[1], Db, Ib)]:
        s = q.double() @ dbt.double().t()
        v, i = s.topk(k, dim=1)
        same = (i == Ig).all(dim=1).float().mean().item()
        res[f"parity_{nm}"] = {"rows_identical": same, "max_abs_D": (v.float() - Dg).abs().max().item()}
        print("parity", nm, res[f"parity_{nm}"], flush=True)
    return res


STAGES = {"gemm": stage_gemm, "small": stage_search_small, "cfg1": stage_cfg1, "time": stage_time}

if __name__ == "__main__":
    want = sys.argv[1:] or list(STAGES)
    os.makedirs("gpurun_out", exist_ok=True)
    for s in want:
        try:
            OUT[s] = STAGES[s]()
        except Exception as e:  # keep going: later stages may still tell us something
            OUT[s] = {"error": repr(e), "trace": traceback.format_exc()}
            print("STAGE FAILED", s, repr(e), flush=True)
            traceback.print_exc()
        with open("gpurun_out/bringup.json", "w") as f:
            json.dump(OUT, f, indent=1, default=str)
    print("done")
