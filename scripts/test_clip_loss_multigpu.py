"""Multi-GPU check and timing of the gathered contrastive loss: every rank runs
keds_b200.contrastive.gathered_clip_loss and the reference's own statement sequence
(src/trainer.py:89-131,164: two dist.all_gather, local-first cat, matmul, two CrossEntropyLoss) in
eager PyTorch on the same features, and compares loss and gradients.

    torchrun --nproc-per-node 2 scripts/test_clip_loss_multigpu.py   -> gpurun_out/clip_loss_<N>gpu.json
"""
from __future__ import annotations

import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keds_b200 import contrastive as kc  # noqa: E402


def reference_loss(image_features, text_features, logit_scale):
    world_size, rank = dist.get_world_size(), dist.get_rank()
    gathered_image_features = [torch.zeros_like(image_features) for _ in range(world_size)]
    gathered_text_features = [torch.zeros_like(text_features) for _ in range(world_size)]
    dist.all_gather(gathered_image_features, image_features)
    dist.all_gather(gathered_text_features, text_features)
    all_image_features = torch.cat([image_features] + gathered_image_features[:rank] + gathered_image_features[rank + 1:])
    all_text_features = torch.cat([text_features] + gathered_text_features[:rank] + gathered_text_features[rank + 1:])
    ground_truth = torch.arange(len(all_image_features)).long().cuda()
    logits_per_image = logit_scale * all_image_features @ all_text_features.t()
    loss_img_val = nn.functional.cross_entropy(logits_per_image, ground_truth)
    loss_txt_val = nn.functional.cross_entropy(logits_per_image.t(), ground_truth)
    return (loss_img_val + loss_txt_val) / 2


def timeit(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main() -> None:
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    torch.backends.cuda.matmul.allow_tf32 = False
    B, d = 128, 768
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    I = torch.nn.functional.normalize(torch.randn(B, d, generator=g, device="cuda"), dim=1)
    T = torch.nn.functional.normalize(I + 6.0 / d ** 0.5 * torch.randn(B, d, generator=g, device="cuda") * d ** 0.5 / d ** 0.5, dim=1)
    out = {}
    grads = {}
    for name, fn in (("native", kc.gathered_clip_loss), ("eager", reference_loss)):
        Ii, Tt = I.clone().requires_grad_(True), T.clone().requires_grad_(True)
        ls = torch.tensor(4.6052, device="cuda", requires_grad=True)   # log(100), model.logit_scale
        loss = fn(Ii, Tt, ls.exp())
        loss.backward()
        grads[name] = (float(loss.detach()), Ii.grad.clone(), Tt.grad.clone(), float(ls.grad))

        def step(fn=fn):
            a, b = I.clone().requires_grad_(True), T.clone().requires_grad_(True)
            s = torch.tensor(4.6052, device="cuda", requires_grad=True)
            fn(a, b, s.exp()).backward()

        out[name + "_ms"] = timeit(step)
    kc.check(torch.cuda.current_device())
    n, e = grads["native"], grads["eager"]
    out.update({
        "world": world, "B": B, "d": d, "loss_native": n[0], "loss_eager": e[0],
        "loss_abs_diff": abs(n[0] - e[0]),
        "dI_rel_diff": float((n[1] - e[1]).abs().max() / e[1].abs().max()),
        "dT_rel_diff": float((n[2] - e[2]).abs().max() / e[2].abs().max()),
        "dlogscale_rel_diff": abs(n[3] - e[3]) / max(abs(e[3]), 1e-9),
    })
    ok = out["loss_abs_diff"] < 1e-4 and out["dI_rel_diff"] < 1e-3 and out["dT_rel_diff"] < 1e-3 and out["dlogscale_rel_diff"] < 1e-2
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    out["all_ranks_ok"] = all(flags)
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/clip_loss_{world}gpu.json", "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out))
    dist.destroy_process_group()
    if not out["all_ranks_ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
