"""Timing of the gallery-ranking metrics at BASELINE.json configs[3] shapes (CIRR 4181 x 2297,
ImageNet-domain 10k x 50k, k up to 200) -> gpurun_out/perf_metrics.json"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keds_b200 import metrics as km  # noqa: E402


def unit(n, d, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(n, d, generator=g, device="cuda")
    return x / x.norm(dim=1, keepdim=True)


def wall(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3


res = {}
# CIRR-shaped
G, Q, d = 2297, 4181, 768
gal = unit(G, d, 1006)
rng = np.random.default_rng(1007)
tgt = rng.integers(0, G, Q)
ref = (tgt + rng.integers(1, G, Q)) % G
qf = gal[torch.from_numpy(tgt).cuda()] + gal[torch.from_numpy(ref).cuda()] + 2.0 * unit(Q, d, 1007)
qf = qf / qf.norm(dim=1, keepdim=True)
index_names = [f"./images/dev/dev-{i}.png" for i in range(G)]
reference_names = [f"dev-{i}.png" for i in ref]
target_names = [f"dev-{i}.png" for i in tgt]
res["cirr_4181x2297_ms"] = wall(lambda: km.get_metrics_cirr(gal, qf, reference_names, index_names, target_names))
res["cirr_rank_kernel_only_ms"] = wall(lambda: km.gallery_rank(qf, gal, tgt, ref))
res["cirr_metrics"] = km.get_metrics_cirr(gal, qf, reference_names, index_names, target_names)
# ImageNet-domain-shaped
NG, NQ = 50000, 10000
glab = torch.from_numpy(rng.integers(0, 7000, NG))
qlab = torch.from_numpy(rng.integers(0, 7000, NQ))
gfe = unit(NG, d, 1008)
qfe = unit(NQ, d, 1009)
res["imgnet_10000x50000_k200_ms"] = wall(lambda: km.get_metrics_imgnet(qfe, gfe, qlab, glab), reps=3)
# COCO-shaped: 5000 pairs, both directions
img = unit(5000, d, 1010)
rf = img + 0.5 * unit(5000, d, 1011)
rf = rf / rf.norm(dim=1, keepdim=True)
res["coco_5000_pairs_ms"] = wall(lambda: km.get_metrics_coco(img, rf, torch.tensor(100.0)), reps=3)
print(json.dumps(res, indent=1, default=float))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/perf_metrics.json", "w"), indent=1, default=float)
