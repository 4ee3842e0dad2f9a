import os, sys, json, torch
sys.path.insert(0, '.')
from keds_b200.index import GpuIndexFlat
g = torch.Generator(device='cuda').manual_seed(1010)
def mk(n):
    x = torch.randn(n, 768, device='cuda', generator=g); return x / x.norm(dim=1, keepdim=True)
N = 1000000
ix = GpuIndexFlat(768, 0, 0)
for i in range(4): ix.add(mk(N // 4))
res = {}
for B, k, Ss in ((4096, 64, (0, 37, 74, 111, 148)), (4096, 16, (0, 19, 37, 74)), (128, 64, (0, 145, 73))):
    q = mk(B)
    for S in Ss:
        if S: os.environ["KEDS_DEBUG_SLICES"] = str(S)
        else: os.environ.pop("KEDS_DEBUG_SLICES", None)
        for _ in range(2): ix.search(q, k)
        ix.sync()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(5): D, I = ix.search(q, k)
        t1.record(); ix.sync()
        st = ix.last_stats()
        res[f"B{B}_k{k}_S{S}"] = {"ms": t0.elapsed_time(t1) / 5, "flagged": st["n_flagged"][0], "slices": st["slices"]}
        print(f"B{B}_k{k}_S{S}", res[f"B{B}_k{k}_S{S}"], flush=True)
json.dump(res, open("gpurun_out/slices.json", "w"), indent=1)
