"""Generate tests/golden/clip_loss.npz (build container only):  python oracle/make_golden_clip_loss.py

The loss of src/trainer.py:85-135,164 is a handful of statements inside get_loss_img2text_image,
whose arithmetic is torch's (matmul, nn.CrossEntropyLoss, autograd). Those statements are run here
as written -- local rows first, then the other ranks' (`:103-112`), `logit_scale * I @ T.t()`
(`:126`), two CrossEntropyLoss calls on the logits and their transpose (`:127-129`), `/ 2` (`:164`)
-- in float64 for two simulated ranks, and the loss and the gradients autograd returns for each
rank's own features and for logit_scale are stored with the inputs.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "clip_loss.npz")


def main() -> None:
    world, B, d = 2, 6, 16
    g = torch.Generator().manual_seed(999)
    feats_i = [torch.nn.functional.normalize(torch.randn(B, d, generator=g, dtype=torch.float64), dim=1) for _ in range(world)]
    feats_t = [torch.nn.functional.normalize(f + 0.3 * torch.randn(B, d, generator=g, dtype=torch.float64), dim=1)
               for f in feats_i]
    loss_img, loss_txt = nn.CrossEntropyLoss(), nn.CrossEntropyLoss()
    out = {"world": np.int64(world), "scale": np.float64(14.285714)}
    for rank in range(world):
        image_features = feats_i[rank].clone().requires_grad_(True)
        text_features = feats_t[rank].clone().requires_grad_(True)
        logit_scale = torch.tensor(14.285714, dtype=torch.float64, requires_grad=True)
        gathered_image_features = [f.clone() for f in feats_i]     # what dist.all_gather fills in (:100)
        gathered_text_features = [f.clone() for f in feats_t]
        all_image_features = torch.cat([image_features] + gathered_image_features[:rank] + gathered_image_features[rank + 1:])
        all_text_features = torch.cat([text_features] + gathered_text_features[:rank] + gathered_text_features[rank + 1:])
        ground_truth = torch.arange(len(all_image_features)).long()
        logits_per_image = logit_scale * all_image_features @ all_text_features.t()
        loss_img_val = loss_img(logits_per_image, ground_truth)
        logits_per_text = logits_per_image.t()
        loss_txt_val = loss_txt(logits_per_text, ground_truth)
        total_loss = (loss_img_val + loss_txt_val) / 2
        total_loss.backward()
        out[f"I{rank}"] = feats_i[rank].numpy()
        out[f"T{rank}"] = feats_t[rank].numpy()
        out[f"loss{rank}"] = total_loss.detach().numpy()
        out[f"dI{rank}"] = image_features.grad.numpy()
        out[f"dT{rank}"] = text_features.grad.numpy()
        out[f"dscale{rank}"] = logit_scale.grad.numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, float(out["loss0"]), float(out["loss1"]))


if __name__ == "__main__":
    main()
