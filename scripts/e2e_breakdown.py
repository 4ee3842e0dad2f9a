"""Where the end-to-end step's overhead over the device-resident step comes from: the same fused
retrieval (128 queries, 2 x 0.5M x 768, k = 16) (a) launched kernel by kernel on a stream,
(b) as a CUDA graph of the four kernels with device-resident queries, (c) the same graph reading
pinned host queries / mirroring (D, I) to pinned host memory, (d) with H2D / D2H copy nodes --
each back to back and with a stream synchronisation per step.  -> gpurun_out/e2e_breakdown.json"""
import json, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keds_b200 import retrieval as kr
from keds_b200.index import METRIC_INNER_PRODUCT, GpuIndexFlat
D, N, B, K = 768, 500_000, 128, 16
dev = torch.device("cuda", 0)
def db(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(n, D, generator=g, device="cuda")
    return x / x.norm(dim=1, keepdim=True)
ia, ib = GpuIndexFlat(D, METRIC_INNER_PRODUCT, 0), GpuIndexFlat(D, METRIC_INNER_PRODUCT, 0)
ia.add(db(N, 1)); ib.add(db(N, 2))
q_dev = db(B, 3); q_host = q_dev.cpu().pin_memory()
perm = torch.randperm(K, generator=torch.Generator().manual_seed(999)).to(dev, torch.int32)
args = dict(topk=K, perm_img=perm, want_feats=True, pool_mode=kr.POOL_SOFTMAX, tau=100.0)
bufs = {}
def stream_step(): kr.retrieve2(ia, ib, q_dev, out=bufs, **args)
# (b) graph over device-resident queries
for _ in range(3): stream_step()
torch.cuda.synchronize()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    stream_step(); side.synchronize()
    g_dev = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_dev, stream=side):
        stream_step()
torch.cuda.current_stream().wait_stream(side)
st_host = kr.RetrievalStep(ia, ib, B, copy_nodes=False, **{k: v for k, v in args.items()})
st_copy = kr.RetrievalStep(ia, ib, B, copy_nodes=True, **{k: v for k, v in args.items()})
for s in (st_host, st_copy): s.q_host.copy_(q_host)
def loop(fn, n, sync):
    cs = torch.cuda.current_stream()
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
        if sync: cs.synchronize()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n * 1e3, 2)
res = []
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for rep in range(3):
    r = {}
    for name, fn in (("a_stream_launches", stream_step), ("b_graph_device_queries", g_dev.replay),
                     ("c_graph_host_io", lambda: st_host.graph.replay()), ("d_graph_copy_nodes", lambda: st_copy.graph.replay())):
        r[name] = {"back_to_back_us": loop(fn, n, False), "sync_per_step_us": loop(fn, n, True)}
    res.append(r); print(json.dumps(r), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/e2e_breakdown.json", "w"), indent=1)
