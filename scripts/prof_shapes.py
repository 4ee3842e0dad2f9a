"""A few searches of one shape, for ncu captures:  python scripts/prof_shapes.py <shape> [iters]
shapes: cfg1 (4096 x 50k, CTA-pair kernel), b4096 (4096 x 0.5M), cfg2 (128 x 2 x 0.5M, retrieve2),
cfg5 (128 x 1M, k = 64), imgnet (10000 x 50k, k = 200), hits (the same shape through
keds_index_label_hits), cirr (gallery rank 4181 x 2297)"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keds_b200 import metrics as km  # noqa: E402
from keds_b200 import retrieval as kr  # noqa: E402
from keds_b200.index import METRIC_INNER_PRODUCT, GpuIndexFlat  # noqa: E402

D = 768


def db(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(n, D, generator=g, device="cuda")
    return x / x.norm(dim=1, keepdim=True)


shape = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
if shape == "cirr":
    gal = db(2297, 1006)
    rng = np.random.default_rng(1007)
    tgt = torch.from_numpy(rng.integers(0, 2297, 4181)).cuda()
    ref = (tgt + 7) % 2297
    q = gal[tgt] + gal[ref] + 2.0 * db(4181, 1007)
    q = q / q.norm(dim=1, keepdim=True)
    for _ in range(iters):
        km.gallery_rank(q, gal, tgt, ref)
elif shape == "hits":
    ix = GpuIndexFlat(D, METRIC_INNER_PRODUCT, 0)
    ix.add(db(50_000, 1000))
    q = db(10_000, 1001)
    rng = np.random.default_rng(1008)
    gl = torch.from_numpy(rng.integers(0, 7000, 50_000)).cuda()
    ql = torch.from_numpy(rng.integers(0, 7000, 10_000)).cuda()
    for _ in range(iters):
        km.index_label_hits(ix, q, gl, ql, [1, 5, 10, 50, 100, 200])
elif shape == "cfg2":
    ia, ib = GpuIndexFlat(D, METRIC_INNER_PRODUCT, 0), GpuIndexFlat(D, METRIC_INNER_PRODUCT, 0)
    ia.add(db(500_000, 1002))
    ib.add(db(500_000, 1003))
    q = db(128, 1004)
    bufs = {}
    for _ in range(iters):
        kr.retrieve2(ia, ib, q, 16, want_feats=True, pool_mode=kr.POOL_SOFTMAX, out=bufs)
else:
    n, b, k = {"cfg1": (50_000, 4096, 16), "b4096": (500_000, 4096, 16), "cfg5": (1_000_000, 128, 64),
               "imgnet": (50_000, 10_000, 200)}[shape]
    ix = GpuIndexFlat(D, METRIC_INNER_PRODUCT, 0)
    ix.add(db(n, 1000))
    q = db(b, 1001)
    for _ in range(iters):
        ix.search(q, k)
torch.cuda.synchronize()
