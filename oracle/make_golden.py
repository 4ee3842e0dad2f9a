"""Generate tests/golden/*.npz|json by running the REFERENCE'S OWN functions, unchanged, from the
read-only tree at /root/reference.  Run in the build container only (the GPU box has no reference):

    python oracle/make_golden.py

The reference modules cannot be imported (src/trainer.py:35-41 and src/eval_utils.py:46-57 import
faiss / llama / webdataset and open data files at import time), so the FunctionDef nodes are cut
out with `ast` and exec'd in a namespace that holds only torch / numpy / os / F.  Nothing from the
reference is copied into this repository: only inputs and the outputs it computed.

Faiss itself is absent, so for the use_faiss=True branch the index objects are stand-ins whose
.search() is the float64 oracle with IndexFlatL2 semantics; that branch therefore pins the
reference's *boundary handling* (query normalisation, numpy float32 in, int64 labels indexing a
CPU tensor, reshape, the batch-shared randperm) -- not Faiss' arithmetic.  The use_faiss=False
branch (src/trainer.py:232-257) and the get_metrics_* functions are pure torch and pin arithmetic.
"""
from __future__ import annotations

import ast
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import knn_oracle as orc  # noqa: E402

REF = os.environ.get("KEDS_REFERENCE", "/root/reference")
OUT = os.environ.get("KEDS_GOLDEN_OUT", os.path.join(ROOT, "tests", "golden"))


def extract(path: str, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "np": np, "os": os, "F": F}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, ns)
    missing = [n for n in names if n not in ns]
    if missing:
        raise RuntimeError(f"{path}: functions not found: {missing}")
    return ns


def unit(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g)
    return x / x.norm(dim=1, keepdim=True)


class StandInIndex:
    """IndexFlatL2 semantics through the float64 oracle (Faiss is not installable here)."""

    def __init__(self, base: np.ndarray):
        self.base = base

    def search(self, x, k):
        assert isinstance(x, np.ndarray) and x.dtype == np.float32 and x.flags["C_CONTIGUOUS"]
        return orc.search(self.base, x, k, "l2")


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    tr = extract(os.path.join(REF, "src", "trainer.py"), ["get_retrieved_features", "get_extra_cap_features"])
    ev = extract(os.path.join(REF, "src", "eval_utils.py"),
                 ["get_metrics_coco", "get_metrics_fashion", "get_metrics_cirr", "get_metrics_imgnet",
                  "get_cirr_testoutput"])

    # ---- retrieval: 2048 x 64 bases (aligned pairs), 32 queries, k = 16
    n, d, b, k = 2048, 64, 32, 16
    image_base = unit(n, d, 1002)
    g = torch.Generator().manual_seed(1003)
    tb = image_base * 0.5 + torch.randn(n, d, generator=g) * 0.5 / np.sqrt(d)
    text_base = tb / tb.norm(dim=1, keepdim=True)
    feature = unit(b, d, 1004) * 3.0  # un-normalised on purpose: the faiss branch normalises (:206)
    basenames = [f"{i:07d}" for i in range(n)]

    # torch branch (use_faiss=False): pure IP matmul + topk + gather, no normalisation, no shuffle
    ti, tt = tr["get_retrieved_features"](feature.clone(), [image_base, text_base], None, topk=k, use_faiss=False)

    # faiss branch with stand-in indices: boundary handling + shared randperm
    database = [image_base, text_base, basenames, StandInIndex(image_base.numpy()), StandInIndex(text_base.numpy())]
    torch.manual_seed(999)
    fi, ft = tr["get_retrieved_features"](feature.clone(), database, None, topk=k, use_faiss=True)
    torch.manual_seed(999)
    perm = torch.randperm(k).numpy()  # the permutation the call above drew (first RNG use after the seed)
    et, en = tr["get_extra_cap_features"](feature.clone(), database, None, topk=2)

    np.savez_compressed(
        os.path.join(OUT, "retrieval.npz"),
        image_base=image_base.numpy(), text_base=text_base.numpy(), feature=feature.numpy(),
        torch_branch_image=ti.numpy(), torch_branch_text=tt.numpy(),
        faiss_branch_image=fi.numpy(), faiss_branch_text=ft.numpy(), perm=perm,
        extra_text=et.numpy(), extra_names=np.array(en),
    )

    # ---- metrics
    metrics = {}
    # CIRR-shaped: names carry a directory so the basename() loop (:1046-1048) matters
    G, Q = 300, 200
    gal = unit(G, d, 1006)
    rng = np.random.default_rng(1007)
    tgt = rng.integers(0, G, Q)
    ref = (tgt + rng.integers(1, G, Q)) % G
    gq = torch.Generator().manual_seed(1007)
    # composed query = target + its reference image + heavy noise: the reference image ranks high
    # (so removing it matters, :1052-1056) and recall lands strictly between 0 and 100
    qf = gal[tgt] + gal[ref] + 3.2 * torch.randn(Q, d, generator=gq) / np.sqrt(d)
    qf = qf / qf.norm(dim=1, keepdim=True)
    index_names = [f"./images/dev/dev-{i}.png" for i in range(G)]
    reference_names = [f"dev-{i}.png" for i in ref]
    target_names = [f"dev-{i}.png" for i in tgt]
    metrics["cirr"] = ev["get_metrics_cirr"](gal, qf, reference_names, index_names, target_names)
    # CIRR test split (src/eval_utils.py:1070-1087): names are matched raw (no basename loop there),
    # the reference image is removed, the first 50 names per pair id are emitted without ".png"
    test_index_names = [f"test1-{i}-img0.png" for i in range(G)]
    test_reference_names = [test_index_names[i] for i in ref]
    pair_ids = torch.arange(7000, 7000 + Q)
    cirr_test = ev["get_cirr_testoutput"](gal, qf, test_reference_names, test_index_names, pair_ids)
    # FashionIQ-shaped
    fnames = [f"B{i:05d}" for i in range(G)]
    metrics["fashion"] = ev["get_metrics_fashion"](gal, qf, fnames, [fnames[i] for i in tgt])
    # COCO-shaped: Q == G pairs
    P = 200
    img = unit(P, d, 1010)
    gc = torch.Generator().manual_seed(1011)
    rf = img + 3.0 * torch.randn(P, d, generator=gc) / np.sqrt(d)
    rf = rf / rf.norm(dim=1, keepdim=True)
    m = ev["get_metrics_coco"](img, rf, torch.tensor(100.0))
    metrics["coco"] = {kk: float(v) for kk, v in m.items()}
    # ImageNet-domain-shaped: labels < 7000 (the function hard-codes num_classes = 7000, :1092)
    NG, NQ, NC = 1500, 250, 40
    glab = torch.from_numpy(rng.integers(0, NC, NG))
    qlab = torch.from_numpy(rng.integers(0, NC, NQ))
    cent = unit(NC, d, 1008)
    gg = torch.Generator().manual_seed(1009)
    gfe = cent[glab] + 1.5 * torch.randn(NG, d, generator=gg) / np.sqrt(d)
    gfe = gfe / gfe.norm(dim=1, keepdim=True)
    qfe = cent[qlab] + 1.5 * torch.randn(NQ, d, generator=gg) / np.sqrt(d)
    qfe = qfe / qfe.norm(dim=1, keepdim=True)
    m = ev["get_metrics_imgnet"](qfe, gfe, qlab, glab)
    metrics["imgnet"] = {kk: float(v) for kk, v in m.items()}

    np.savez_compressed(
        os.path.join(OUT, "metrics_inputs.npz"),
        gal=gal.numpy(), qf=qf.numpy(), tgt=tgt, ref=ref,
        coco_img=img.numpy(), coco_ref=rf.numpy(),
        in_gal=gfe.numpy(), in_q=qfe.numpy(), in_glab=glab.numpy(), in_qlab=qlab.numpy(),
    )
    with open(os.path.join(OUT, "metrics_expected.json"), "w") as f:
        json.dump({"index_names": index_names, "reference_names": reference_names,
                   "target_names": target_names, "fashion_names": fnames, "metrics": metrics,
                   "cirr_test": {"index_names": test_index_names, "reference_names": test_reference_names,
                                 "pair_ids": pair_ids.tolist(), "output": cirr_test}}, f, indent=1)
    print("wrote", sorted(os.listdir(OUT)))
    for kk, v in metrics.items():
        print(kk, v)


if __name__ == "__main__":
    main()
