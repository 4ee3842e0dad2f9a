"""Generate tests/golden/consumer_train.npz by running the REFERENCE'S OWN module classes in TRAIN
mode, unchanged, from the read-only tree at /root/reference (build container only):

    python oracle/make_golden_consumer_train.py

Same extraction as make_golden_consumer.py (ClassDef nodes of CrossAttention, CrossFormer, IM2TEXT,
src/model/model.py:37-123, exec'd with torch / nn / einops only). The modules are built in float64
with dropout 0 (so that train mode is deterministic), run through the call sequence of
src/trainer.py:59-69, and differentiated with torch.autograd against a fixed upstream gradient: the
fixture holds weights, inputs, tokens and the gradient of every parameter.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from make_golden_consumer import REF, ROOT, extract_classes

OUT = os.path.join(ROOT, "tests", "golden", "consumer_train.npz")


def main() -> None:
    ns = extract_classes(os.path.join(REF, "src", "model", "model.py"), ["CrossAttention", "CrossFormer", "IM2TEXT"])
    d_in, d_mid, d_tok, n_hidden, n_layers, heads, dim_head = 48, 32, 40, 2, 3, 4, 8
    B, k = 6, 5
    torch.manual_seed(4321)
    img2text = ns["IM2TEXT"](embed_dim=d_in, middle_dim=d_mid, output_dim=d_tok, n_layer=n_hidden, dropout=0.0).double().train()
    fuse = ns["CrossFormer"](q_dim=d_tok, k_dim=d_tok, v_dim=d_tok, num_layers=n_layers, heads=heads,
                             dim_head=dim_head).double().train()
    cond = ns["CrossFormer"](q_dim=d_tok, k_dim=d_tok, v_dim=d_tok, num_layers=n_layers, heads=heads,
                             dim_head=dim_head).double().train()
    g = torch.Generator().manual_seed(77)
    feat = torch.randn(B, d_in, generator=g, dtype=torch.float64)
    topk_image = torch.randn(B, k, d_in, generator=g, dtype=torch.float64)
    topk_text = torch.randn(B, k, d_in, generator=g, dtype=torch.float64)
    dtokens = torch.randn(B, 3, d_tok, generator=g, dtype=torch.float64)
    # the call sequence of src/trainer.py:59-69
    mapped = img2text(feat)
    nb_img = img2text(topk_image)
    nb_txt = img2text(topk_text)
    fused = fuse(mapped.unsqueeze(1), nb_img, nb_img)
    text_c = cond(mapped.unsqueeze(1), nb_txt, nb_txt)
    tokens = torch.cat([fused, text_c, mapped.unsqueeze(1)], dim=1)
    tokens.backward(dtokens)
    out = {
        "dims": np.array([d_in, d_mid, d_tok, n_hidden, n_layers, heads, dim_head], dtype=np.int64),
        "feat": feat.numpy(), "topk_image": topk_image.numpy(), "topk_text": topk_text.numpy(),
        "dtokens": dtokens.numpy(), "tokens": tokens.detach().numpy(),
    }
    for prefix, mod in (("img2text", img2text), ("retrieval_fuse", fuse), ("text_condition", cond)):
        for name, t in mod.named_parameters():
            out[f"{prefix}/{name}"] = t.detach().numpy()
            out[f"grad/{prefix}/{name}"] = t.grad.numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, len(out), "arrays")


if __name__ == "__main__":
    main()
