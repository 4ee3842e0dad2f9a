"""Generate tests/golden/consumer.npz by running the REFERENCE'S OWN module classes, unchanged,
from the read-only tree at /root/reference (build container only):

    python oracle/make_golden_consumer.py

src/model/model.py imports open_clip / llama pieces at module level, so the ClassDef nodes of
CrossAttention, CrossFormer and IM2TEXT (src/model/model.py:37-123) are cut out with `ast` and
exec'd in a namespace holding torch / nn / einops only. Nothing from the reference is copied into
this repository: the fixture holds the randomly initialised weights (small widths), the inputs and
the outputs the reference modules computed for the call sequence of src/trainer.py:59-69 /
src/eval_utils.py:378-383 in eval mode.
"""
from __future__ import annotations

import ast
import os

import numpy as np
import torch
import torch.nn as nn
from einops import rearrange, repeat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("KEDS_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "consumer.npz")


def extract_classes(path: str, names):
    tree = ast.parse(open(path).read())
    ns = {"torch": torch, "nn": nn, "rearrange": rearrange, "repeat": repeat}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    if missing:
        raise RuntimeError(f"{path}: classes not found: {missing}")
    return ns


def main() -> None:
    ns = extract_classes(os.path.join(REF, "src", "model", "model.py"), ["CrossAttention", "CrossFormer", "IM2TEXT"])
    d_in, d_mid, d_tok, n_hidden, n_layers, heads, dim_head = 48, 32, 40, 2, 3, 4, 8
    B, k, n_base = 5, 6, 64
    torch.manual_seed(999)
    img2text = ns["IM2TEXT"](embed_dim=d_in, middle_dim=d_mid, output_dim=d_tok, n_layer=n_hidden).eval()
    fuse = ns["CrossFormer"](q_dim=d_tok, k_dim=d_tok, v_dim=d_tok, num_layers=n_layers, heads=heads,
                             dim_head=dim_head).eval()
    cond = ns["CrossFormer"](q_dim=d_tok, k_dim=d_tok, v_dim=d_tok, num_layers=n_layers, heads=heads,
                             dim_head=dim_head).eval()
    g = torch.Generator().manual_seed(1234)
    feat = torch.randn(B, d_in, generator=g)
    base_img = torch.randn(n_base, d_in, generator=g)
    base_txt = torch.randn(n_base, d_in, generator=g)
    I_img = torch.randint(0, n_base, (B, k), generator=g)
    I_txt = torch.randint(0, n_base, (B, k), generator=g)
    with torch.no_grad():
        topk_image = base_img[I_img.reshape(-1)].reshape(B, k, -1)
        topk_text = base_txt[I_txt.reshape(-1)].reshape(B, k, -1)
        # the call sequence of src/trainer.py:59-69
        mapped = img2text(feat)
        nb_img = img2text(topk_image)
        nb_txt = img2text(topk_text)
        fused = fuse(mapped.unsqueeze(1), nb_img, nb_img)
        text_c = cond(mapped.unsqueeze(1), nb_txt, nb_txt)
        tokens = torch.cat([fused, text_c, mapped.unsqueeze(1)], dim=1)
    out = {
        "dims": np.array([d_in, d_mid, d_tok, n_hidden, n_layers, heads, dim_head], dtype=np.int64),
        "feat": feat.numpy(), "base_img": base_img.numpy(), "base_txt": base_txt.numpy(),
        "I_img": I_img.numpy(), "I_txt": I_txt.numpy(), "tokens": tokens.numpy(),
    }
    for prefix, mod in (("img2text", img2text), ("retrieval_fuse", fuse), ("text_condition", cond)):
        for name, t in mod.state_dict().items():
            out[f"{prefix}/{name}"] = t.numpy()
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items() if "/" not in k})


if __name__ == "__main__":
    main()
