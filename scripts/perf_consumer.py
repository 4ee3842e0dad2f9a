"""Timing of the neighbour consumer at the training-step shape (B = 128, k = 16,
768 -> 512 -> 512 -> 768 MLP, 2 x 3 cross-attention layers of 8 x 64) against the same modules in
eager PyTorch on the same GPU (gather included on both sides).  -> gpurun_out/perf_consumer.json

    python scripts/perf_consumer.py [B] [k]
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keds_b200.consumer import NeighbourConsumer  # noqa: E402
from keds_b200.index import METRIC_INNER_PRODUCT, GpuIndexFlat  # noqa: E402
from oracle import consumer_oracle as corc  # noqa: E402  (seeded weights + the float64 check)


def timeit(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def eager_tokens(sds, heads, feat, base_img, base_txt, I_img, I_txt):
    """The reference's op sequence (src/model/model.py:56-79,98-101,120-123; src/trainer.py:59-69)."""
    m, fs, ts = sds

    def im2text(x):
        i = 0
        while f"layers.{i}.0.weight" in m:
            x = F.relu(F.linear(x, m[f"layers.{i}.0.weight"], m[f"layers.{i}.0.bias"]))
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
    sds = tuple({kk: torch.from_numpy(v).cuda() for kk, v in sd.items()} for sd in sds_np)
    g = torch.Generator(device="cuda").manual_seed(5)
    base_img = F.normalize(torch.randn(n, 768, generator=g, device="cuda"), dim=1)
    base_txt = F.normalize(torch.randn(n, 768, generator=g, device="cuda"), dim=1)
    feat = F.normalize(torch.randn(B, 768, generator=g, device="cuda"), dim=1)
    I_img = torch.randint(0, n, (B, k), generator=g, device="cuda")
    I_txt = torch.randint(0, n, (B, k), generator=g, device="cuda")
    ix_i, ix_t = GpuIndexFlat(768, METRIC_INNER_PRODUCT, 0), GpuIndexFlat(768, METRIC_INNER_PRODUCT, 0)
    ix_i.add(base_img)
    ix_t.add(base_txt)
    cons = NeighbourConsumer(*sds, heads=8, device=0)
    out = torch.empty(B, 3, 768, device="cuda")

    got = cons(feat, ix_i, ix_t, I_img, I_txt, out=out).cpu().numpy()
    launches_per_call = cons.check()
    want = corc.consumer_tokens(*sds_np, 8, feat.cpu().numpy(), base_img.cpu().numpy(), base_txt.cpu().numpy(),
                                I_img.cpu().numpy(), I_txt.cpu().numpy())
    res = {"B": B, "k": k, "launches_per_call": launches_per_call,
           "rel_err_native_vs_f64": float(np.abs(got - want).max() / np.abs(want).max())}
    res["native_ms"] = timeit(lambda: cons(feat, ix_i, ix_t, I_img, I_txt, out=out))
    cons.check()
    # the same call sequence replayed from a CUDA graph (no host launch cost)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        cons(feat, ix_i, ix_t, I_img, I_txt, out=out)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        cons(feat, ix_i, ix_t, I_img, I_txt, out=out)
    res["native_graph_ms"] = timeit(graph.replay)
    cons.check()
    # per-CTA stage times of every k_linear_tf32 launch (in-kernel %globaltimer), warm
    cons.set_debug(True)
    cons(feat, ix_i, ix_t, I_img, I_txt, out=out)
    cons(feat, ix_i, ix_t, I_img, I_txt, out=out)
    M33, Bk = B * (1 + 2 * k), B * k
    mt = lambda m: (m + 127) // 128
    sms = torch.cuda.get_device_properties(0).multi_processor_count

    def ctas(m, n, nz):  # the host's tile-width rule (consumer_host.cuh: consumer_linear)
        c128, c256 = mt(m) * ((n + 127) // 128) * nz, mt(m) * ((n + 255) // 256) * nz
        if c128 * 6 <= sms:  # split-K clusters: 4 CTAs per 32-column tile
            tiles = mt(m) * ((n + 31) // 32) * nz
            return (4 if tiles * 4 <= sms else 2) * tiles
        return c256 if -(-c256 // sms) * 48 < -(-c128 // sms) * 32 else c128

    grids = [ctas(M33, 512, 1), ctas(M33, 512, 1), ctas(M33, 768, 1), ctas(Bk, 3072, 2)] \
        + [ctas(B, 512, 2), ctas(B, 768, 2)] * 3
    tl = []
    t_first = None
    for i, gsz in enumerate(grids):
        if gsz > 1024:
            tl.append({"launch": i, "ctas": gsz, "skipped": "more CTAs than the debug buffer holds"})
            continue
        t = cons.debug_timeline(i, gsz).astype(np.int64)
        if not t[:, 0].any():
            tl.append({"launch": i, "ctas": gsz, "skipped": "persistent kernel: no per-CTA stamps"})
            continue
        t_first = int(t[:, 0].min()) if t_first is None else t_first
        if i == 3:  # the key/value projection: per-CTA (start, end) relative to the launch's first CTA
            res["kv_cta_us"] = ((t[:, [0, 4]] - t[:, 0].min()) / 1e3).round(2).tolist()
        d = np.diff(t, axis=1)
        tl.append({"launch": i, "ctas": gsz, "span_us": float(t[:, 4].max() - t[:, 0].min()) / 1e3,
                   "start_us": float(t[:, 0].min() - t_first) / 1e3, "end_us": float(t[:, 4].max() - t_first) / 1e3,
                   "mean_us": {n: float(d[:, j].mean()) / 1e3 for j, n in
                               enumerate(("prologue", "dependency_wait", "mainloop", "epilogue"))}})
    res["linear_timeline"] = tl
    cons.set_debug(False)
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        with torch.no_grad():
            e = eager_tokens(sds, 8, feat, base_img, base_txt, I_img, I_txt).cpu().numpy()
            key = "eager_tf32" if tf32 else "eager_fp32"
            res[key + "_rel_err_vs_f64"] = float(np.abs(e - want).max() / np.abs(want).max())
            res[key + "_ms"] = timeit(lambda: eager_tokens(sds, 8, feat, base_img, base_txt, I_img, I_txt))
    M = B * (1 + 2 * k)
    flops = 2.0 * M * (768 * 512 + 512 * 512 + 512 * 768) + 2.0 * (2 * B * k) * 768 * (3 * 2 * 512) \
        + 2 * 3 * 2.0 * B * (768 * 512 + 512 * 768)
    res["gflop"] = flops / 1e9
    res["native_tflops"] = flops / (res["native_ms"] * 1e-3) / 1e12
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/perf_consumer.json", "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
