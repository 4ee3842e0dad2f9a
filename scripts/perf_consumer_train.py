"""Timing of the neighbour consumer in TRAINING mode (forward + backward of src/trainer.py:59-69 at
B = 128, k = 16, the reference's widths, dropout 0.1) against the same op sequence in eager PyTorch
autograd on the same GPU (fp32, and with TF32 matmuls allowed).  -> gpurun_out/perf_consumer_train.json

    python scripts/perf_consumer_train.py [B] [k]
"""
from __future__ import annotations

import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keds_b200.consumer import TrainableNeighbourConsumer  # noqa: E402
from keds_b200.index import METRIC_INNER_PRODUCT, GpuIndexFlat  # noqa: E402
from oracle import consumer_oracle as corc  # noqa: E402  (seeded weights only)
from perf_consumer import timeit  # noqa: E402


def eager_tokens(sds, heads, feat, base_img, base_txt, I_img, I_txt, p_drop):
    m, fs, ts = sds

    def im2text(x):
        i = 0
        while f"layers.{i}.0.weight" in m:
            x = F.relu(F.dropout(F.linear(x, m[f"layers.{i}.0.weight"], m[f"layers.{i}.0.bias"]), p_drop, True))
            i += 1
        return F.linear(x, m["fc_out.weight"], m["fc_out.bias"])

    def former(sd, q, kv):
        l = 0
        while f"cross_layers.{l}.to_q.weight" in sd:
            p = f"cross_layers.{l}."
            B, n, _ = kv.shape
            Q = F.linear(q, sd[p + "to_q.weight"], sd[p + "to_q.bias"]).view(B, -1, heads, 64).transpose(1, 2)
            K = F.linear(kv, sd[p + "to_k.weight"], sd[p + "to_k.bias"]).view(B, n, heads, 64).transpose(1, 2)
            V = F.linear(kv, sd[p + "to_v.weight"], sd[p + "to_v.bias"]).view(B, n, heads, 64).transpose(1, 2)
            dots = torch.einsum("bhid,bhjd->bhij", Q, K) * 64 ** -0.5
            out = torch.einsum("bhij,bhjd->bhid", dots.softmax(dim=-1), V).transpose(1, 2).reshape(B, -1, heads * 64)
            q = F.linear(out, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])
            l += 1
        return q

    B, k = I_img.shape
    mapped = im2text(feat)
    nb_img = im2text(base_img[I_img.reshape(-1)].reshape(B, k, -1))
    nb_txt = im2text(base_txt[I_txt.reshape(-1)].reshape(B, k, -1))
    fused = former(fs, mapped.unsqueeze(1), nb_img)
    text_c = former(ts, mapped.unsqueeze(1), nb_txt)
    return torch.cat([fused, text_c, mapped.unsqueeze(1)], dim=1)


def main() -> None:
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    n = 100_000
    sds_np = corc.random_state_dicts(768, 512, 768, 2, 3, 8, 64, seed=11)
    g = torch.Generator(device="cuda").manual_seed(5)
    base_img = F.normalize(torch.randn(n, 768, generator=g, device="cuda"), dim=1)
    base_txt = F.normalize(torch.randn(n, 768, generator=g, device="cuda"), dim=1)
    feat = F.normalize(torch.randn(B, 768, generator=g, device="cuda"), dim=1)
    I_img = torch.randint(0, n, (B, k), generator=g, device="cuda")
    I_txt = torch.randint(0, n, (B, k), generator=g, device="cuda")
    dtok = torch.randn(B, 3, 768, generator=g, device="cuda")
    ia, ib = GpuIndexFlat(768, METRIC_INNER_PRODUCT, 0), GpuIndexFlat(768, METRIC_INNER_PRODUCT, 0)
    ia.add(base_img)
    ib.add(base_txt)
    res = {"B": B, "k": k}

    mod = TrainableNeighbourConsumer(*({kk: torch.from_numpy(v) for kk, v in sd.items()} for sd in sds_np), heads=8,
                                     device=0, dropout=0.1).train()

    def native_step():
        mod.flat.grad = None
        mod(feat, ia, ib, I_img, I_txt).backward(dtok)

    n0 = mod.check()
    native_step()
    res["native_launches_per_step"] = mod.check() - n0
    res["native_fwd_bwd_ms"] = timeit(native_step, 100)

    def native_fwd():
        with torch.no_grad():
            mod(feat, ia, ib, I_img, I_txt)

    res["native_fwd_only_ms"] = timeit(native_fwd, 100)

    # the same module under torch.cuda.make_graphed_callables: forward and backward replayed as two
    # CUDA graphs (the launch sequence is what bounds the native step, not its kernels)
    try:
        class Step(torch.nn.Module):
            def __init__(self, m):
                super().__init__()
                self.m = m

            def forward(self, f, ii, it):
                return self.m(f, ia, ib, ii, it)

        gmod = TrainableNeighbourConsumer(*({kk: torch.from_numpy(v) for kk, v in sd.items()} for sd in sds_np), heads=8,
                                          device=0, dropout=0.1).train()
        f_req = feat.clone().requires_grad_(False)
        graphed = torch.cuda.make_graphed_callables(Step(gmod), (f_req, I_img, I_txt))

        def graphed_step():
            gmod.flat.grad = None
            graphed(f_req, I_img, I_txt).backward(dtok)

        graphed_step()
        torch.cuda.synchronize()
        res["graphed_fwd_bwd_ms"] = timeit(graphed_step, 100)
        # same parameters, dropout off: the graphed step must reproduce the plain native step
        gmod.eval(); mod.eval()
    except Exception as e:  # reported, not fatal
        res["graphed_error"] = repr(e)[:400]

    params = tuple({kk: torch.from_numpy(v).cuda().requires_grad_(True) for kk, v in sd.items()} for sd in sds_np)

    def eager_step():
        for sd in params:
            for t in sd.values():
                t.grad = None
        eager_tokens(params, 8, feat, base_img, base_txt, I_img, I_txt, 0.1).backward(dtok)

    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        res[f"eager_fwd_bwd_ms_{'tf32' if tf32 else 'fp32'}"] = timeit(eager_step, 30)
    res["speedup_vs_eager_fp32"] = res["eager_fwd_bwd_ms_fp32"] / res["native_fwd_bwd_ms"]
    res["speedup_vs_eager_tf32"] = res["eager_fwd_bwd_ms_tf32"] / res["native_fwd_bwd_ms"]
    # forward 29.4 GFLOP at B = 128, k = 16; backward twice that
    flop = 3 * (2.0 * B * (1 + 2 * k) * (768 * 512 + 512 * 512 + 512 * 768) + 2.0 * 2 * B * k * 768 * 3072)
    res["tflops_native"] = flop / (res["native_fwd_bwd_ms"] * 1e-3) / 1e12
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/perf_consumer_train.json", "w"), indent=1)


if __name__ == "__main__":
    main()
