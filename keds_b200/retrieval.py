"""The retrieval operator of KEDs on the native path.

`get_retrieved_features` / `get_extra_cap_features` keep the reference's names, arguments and
return values (src/trainer.py:198-283, src/eval_utils.py:153-186) so the training and eval loops
call them unchanged.  What changes underneath:

  reference (per call)                                   here
  -----------------------------------------------------  -----------------------------------------
  feature.clone().cpu().numpy()  x2  (D2H + sync)        queries stay on the GPU
  image_index.search + text_index.search (2 passes)      one fused pass over both databases
  base[I.reshape(-1)] on a 1.5 GB CPU tensor             gather kernel on the resident fp32 rows
  feats[:, randperm(k), :] CPU copy                      permutation folded into the gather
  .clone().to(device)  x2 (12.6 MB pageable H2D)         output is produced on the device

`KnowledgeBase` replaces the index construction block of src/main.py:72-96 /
src/eval_retrieval.py:280-298: it loads the .pt databases (layout unchanged: float32 [N, 768] CPU
tensors + database_names.txt) and builds one native index per database on the local GPU.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _capi
from .index import GpuIndexFlat, METRIC_INNER_PRODUCT, METRIC_L2, search2, _stream_ptr


def gather_rows(index: GpuIndexFlat, I: torch.Tensor, perm: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[b, j, :] = rows[I[b, perm[j]], :] from the index's resident fp32 rows (device tensors)."""
    lib = _capi.load()
    assert I.is_cuda and I.dtype == torch.int64 and I.dim() == 2
    I = I.contiguous()
    B, k = I.shape
    out = torch.empty((B, k, index.d), dtype=torch.float32, device=I.device)
    p = 0
    if perm is not None:
        perm = perm.to(device=I.device, dtype=torch.int32).contiguous()
        assert perm.numel() == k
        p = perm.data_ptr()
    _capi.check(
        lib.keds_gather_pool(index.rows_ptr(), index.ntotal, I.data_ptr(), 0, p, B, k, 1, index.d,
                             out.data_ptr(), _stream_ptr(index.device))
    )
    return out


def weighted_pool(index: GpuIndexFlat, I: torch.Tensor, W: torch.Tensor) -> torch.Tensor:
    """out[b, h, :] = sum_j W[b, h, j] * rows[I[b, j], :]   (W: [B, H, k] float32, device)."""
    lib = _capi.load()
    assert I.is_cuda and I.dtype == torch.int64 and W.is_cuda and W.dtype == torch.float32
    I = I.contiguous()
    W = W.contiguous()
    B, k = I.shape
    assert W.dim() == 3 and W.shape[0] == B and W.shape[2] == k
    H = W.shape[1]
    out = torch.empty((B, H, index.d), dtype=torch.float32, device=I.device)
    _capi.check(
        lib.keds_gather_pool(index.rows_ptr(), index.ntotal, I.data_ptr(), W.data_ptr(), 0, B, k, H,
                             index.d, out.data_ptr(), _stream_ptr(index.device))
    )
    return out


def search_gather(index: GpuIndexFlat, q: torch.Tensor, k: int, bases: Optional[GpuIndexFlat] = None,
                  weights: Optional[torch.Tensor] = None, perm: Optional[torch.Tensor] = None):
    """Search, then consume the neighbours on the device (the extension SURVEY.md section 8(b) names
    next to the Faiss surface): returns (D, I, out) with
        out[b, j, :] = rows[I[b, perm[j]], :]                 weights is None   ([B, k, d])
        out[b, h, :] = sum_j weights[b, h, j] rows[I[b, j], :]   weights [B, H, k]  ([B, H, d])
    `rows` are the resident fp32 rows of `bases` (an index aligned row for row with `index`, e.g. the
    text database for an image search -- src/trainer.py:214-216 gathers both bases with one label
    set) or of `index` itself. q: CUDA float32 [B, d]."""
    q = index._check_q_tensor(q)
    D, I = index.search(q, int(k))
    src = bases if bases is not None else index
    if src.ntotal != index.ntotal:
        raise ValueError("search_gather: `bases` must be aligned row for row with the searched index")
    out = weighted_pool(src, I, weights) if weights is not None else gather_rows(src, I, perm)
    return D, I, out


class KnowledgeBase:
    """The `database` object of the reference as a sequence:
    [image_bases, text_bases, basenames, image_gpu_index, text_gpu_index]  (src/main.py:95-96).
    image_bases / text_bases stay plain CPU float32 tensors (the .pt layout); the indices are
    native and hold their own device copies."""

    def __init__(self, image_bases: torch.Tensor, text_bases: torch.Tensor, basenames: Sequence[str],
                 device: int = 0, metric: int = METRIC_L2) -> None:
        self._init_indices(image_bases, text_bases, basenames, device, metric, normalize=False)

    def _init_indices(self, image_bases, text_bases, basenames, device, metric, normalize) -> None:
        for t in (image_bases, text_bases):
            if not (isinstance(t, torch.Tensor) and t.dtype == torch.float32 and t.dim() == 2):
                raise TypeError("knowledge bases must be 2-D float32 torch tensors (the .pt layout)")
        if image_bases.shape != text_bases.shape:
            raise ValueError("image and text bases must be aligned row for row")
        if len(basenames) != image_bases.shape[0]:
            raise ValueError("one basename per row")
        self.basenames = list(basenames)
        d = image_bases.shape[1]
        self.image_index = GpuIndexFlat(d, metric, device)
        self.text_index = GpuIndexFlat(d, metric, device)
        self.image_index.add(image_bases, normalize=normalize)
        self.text_index.add(text_bases, normalize=normalize)
        if normalize:  # keep the host tensors bit-identical to what the device searches
            image_bases = torch.from_numpy(self.image_index.get_rows())
            text_bases = torch.from_numpy(self.text_index.get_rows())
        self.image_bases = image_bases
        self.text_bases = text_bases

    @classmethod
    def load(cls, image_pt: str, text_pt: str, names_txt: str, device: int = 0, metric: int = METRIC_L2):
        """torch.load the two .pt files and read database_names.txt (src/main.py:470-477)."""
        image_bases = torch.load(image_pt, map_location="cpu")
        text_bases = torch.load(text_pt, map_location="cpu")
        with open(names_txt) as f:
            basenames = [line.strip() for line in f]  # line.strip(), as src/main.py:474-475
        return cls(image_bases, text_bases, basenames, device, metric)

    # sequence protocol: database[0..4]
    def _seq(self):
        return [self.image_bases, self.text_bases, self.basenames, self.image_index, self.text_index]

    def __getitem__(self, i):
        return self._seq()[i]

    def __len__(self) -> int:
        return 5


def _native(ix) -> bool:
    return isinstance(ix, GpuIndexFlat)


def get_retrieved_features(feature, database, args=None, topk: int = 16, use_faiss: bool = True):
    """Same contract as src/trainer.py:198-259: returns (topk_image_features, topk_text_features),
    each [B, topk, d] on feature.device; image neighbours permuted along k by one
    torch.randperm(topk) shared by the batch (:218-219), text neighbours in rank order.

    use_faiss=False in the reference is a dense torch matmul over the whole database; both
    settings run the native search here (same answer: exact top-k by inner product / L2)."""
    image_base, text_base, basenames, image_index, text_index = (database[i] for i in range(5))
    if not (_native(image_index) and _native(text_index)):
        raise TypeError("database[3] and database[4] must be keds_b200 GPU indices (no CPU path)")
    if use_faiss:
        feature = feature / feature.norm(dim=1, keepdim=True)  # src/trainer.py:206
    dev = torch.device("cuda", image_index.device)
    q = feature.detach().to(device=dev, dtype=torch.float32).contiguous()
    flags = 0 if use_faiss else _capi.SEARCH_FORCE_IP  # the torch branch ranks by raw inner product
    perm = torch.randperm(topk) if use_faiss else None  # CPU generator, like the reference (:218)
    out = retrieve2(image_index, text_index, q, topk, perm_img=perm, want_feats=True, pool_mode=0, flags=flags)
    return out["feat_img"].to(feature.device), out["feat_txt"].to(feature.device)


POOL_NONE, POOL_MEAN, POOL_SOFTMAX = 0, 1, 2


def retrieve2(image_index: GpuIndexFlat, text_index: GpuIndexFlat, q: torch.Tensor, topk: int = 16,
              perm_img: Optional[torch.Tensor] = None, perm_txt: Optional[torch.Tensor] = None,
              want_feats: bool = True, pool_mode: int = POOL_NONE, tau: float = 100.0, flags: int = 0,
              out: Optional[dict] = None, host_out: Optional[dict] = None) -> dict:
    """One native call for the whole operator: fused two-database search, then for each stream the
    gathered neighbours [B, k, d] (image stream permuted by perm_img) and/or the pooled stream
    [B, d].  `out` may carry preallocated tensors from a previous call (keys as returned) so a
    steady-state loop allocates nothing.

    Host I/O without copies (keds_retrieve2_hostio): `q` may be a pinned CPU tensor -- the first
    kernel reads it through the device mapping -- and `host_out` may hold pinned CPU tensors
    D_img / I_img / D_txt / I_txt ([B, k]) that the kernels fill next to the device results; they
    are complete once the stream has been synchronised."""
    lib = _capi.load()
    hostio = host_out is not None or not q.is_cuda
    if q.is_cuda:
        q = image_index._check_q_tensor(q)
        dev = q.device
    else:
        if not q.is_pinned() or q.dtype != torch.float32 or q.dim() != 2 or q.shape[1] != image_index.d or not q.is_contiguous():
            raise TypeError("retrieve2: a host query must be a pinned, contiguous float32 [B, d] tensor")
        dev = torch.device("cuda", image_index.device)
    B, d, k = q.shape[0], image_index.d, int(topk)
    o = out if out is not None else {}

    def buf(name, shape, dtype):
        t = o.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != dev:
            t = torch.empty(shape, dtype=dtype, device=dev)
            o[name] = t
        return t

    Di, Ii = buf("D_img", (B, k), torch.float32), buf("I_img", (B, k), torch.int64)
    Dt, It = buf("D_txt", (B, k), torch.float32), buf("I_txt", (B, k), torch.int64)
    fi = ft = pi = pt = None
    if want_feats:
        fi, ft = buf("feat_img", (B, k, d), torch.float32), buf("feat_txt", (B, k, d), torch.float32)
    if pool_mode != POOL_NONE:
        pi, pt = buf("pool_img", (B, d), torch.float32), buf("pool_txt", (B, d), torch.float32)

    def perm_ptr(p):
        if p is None:
            return 0, None
        p = p.to(device=dev, dtype=torch.int32).contiguous()
        assert p.numel() == k
        return p.data_ptr(), p

    pi_ptr, _keep1 = perm_ptr(perm_img)
    pt_ptr, _keep2 = perm_ptr(perm_txt)
    ptr = lambda t: 0 if t is None else t.data_ptr()
    if hostio:
        ho = host_out or {}
        hp = []
        for name, dt in (("D_img", torch.float32), ("I_img", torch.int64), ("D_txt", torch.float32), ("I_txt", torch.int64)):
            t = ho.get(name)
            if t is not None and not (t.is_pinned() and t.dtype == dt and t.is_contiguous() and t.numel() == B * k):
                raise TypeError(f"retrieve2: host_out[{name!r}] must be a pinned contiguous {dt} tensor of {B * k} elements")
            hp.append(0 if t is None else t.data_ptr())
        _capi.check(
            lib.keds_retrieve2_hostio(image_index._h, text_index._h, q.data_ptr(), B, k, pi_ptr, pt_ptr, int(pool_mode),
                                      float(tau), Di.data_ptr(), Ii.data_ptr(), Dt.data_ptr(), It.data_ptr(),
                                      hp[0], hp[1], hp[2], hp[3], ptr(fi), ptr(ft), ptr(pi), ptr(pt), int(flags),
                                      _stream_ptr(image_index.device))
        )
        return o
    _capi.check(
        lib.keds_retrieve2(image_index._h, text_index._h, q.data_ptr(), B, k, pi_ptr, pt_ptr, int(pool_mode),
                           float(tau), Di.data_ptr(), Ii.data_ptr(), Dt.data_ptr(), It.data_ptr(), ptr(fi),
                           ptr(ft), ptr(pi), ptr(pt), int(flags), _stream_ptr(image_index.device))
    )
    return o


class RetrievalStep:
    """The steady-state retrieval step of a training loop, captured once into a CUDA graph:

        pinned host queries --H2D--> fused two-database search + gather + pool --D2H--> (D, I)

    One graph launch per step replaces ~10 stream operations, so the host cost of a step is a
    few microseconds. Write the batch into `q_host` (pinned, [batch, d]) or pass it to `run()`;
    after `run()` returns, `D_img/I_img/D_txt/I_txt` (pinned host) hold what `index.search` gives
    the host in the reference, and `out` holds the device tensors (gathered / pooled streams).
    """

    def __init__(self, image_index: GpuIndexFlat, text_index: GpuIndexFlat, batch: int, topk: int = 16,
                 perm_img: Optional[torch.Tensor] = None, want_feats: bool = True,
                 pool_mode: int = POOL_NONE, tau: float = 100.0, copy_nodes: Optional[bool] = None) -> None:
        self.ia, self.ib = image_index, text_index
        # copy_nodes=False: no copies in the graph -- the first kernel reads the pinned queries through
        # the mapping and the ranking blocks store (D, I) into the pinned result block next to the
        # device one (keds_retrieve2_hostio). True: H2D + D2H copy nodes around the search (the
        # round-1 layout). None (default): capture both once and keep the faster on THIS box -- SM
        # loads over PCIe beat the copy engines' set-up cost on most boxes (7 of 8 measured), but how
        # the pinned pages sit relative to the GPU's root complex decides it; the answers are the same.
        # KEDS_STEP_COPY_NODES=0|1 pins the choice.
        env = os.environ.get("KEDS_STEP_COPY_NODES")
        if copy_nodes is None and env in ("0", "1"):
            copy_nodes = env == "1"
        self._auto_layout = copy_nodes is None
        self.copy_nodes = bool(copy_nodes)
        self.layout_probe_us = None
        dev = torch.device("cuda", image_index.device)
        d, k = image_index.d, int(topk)
        self.q_host = torch.empty((batch, d), dtype=torch.float32).pin_memory()
        self.q_dev = torch.empty((batch, d), dtype=torch.float32, device=dev)
        # (I_img | I_txt | D_img | D_txt) live in one device block mirrored by one pinned host block,
        # so the results leave the GPU with a single copy
        nI, nD = batch * k * 8, batch * k * 4
        self._res_dev = torch.empty(2 * nI + 2 * nD, dtype=torch.uint8, device=dev)
        self._res_host = torch.empty(2 * nI + 2 * nD, dtype=torch.uint8).pin_memory()

        def views(blk):
            return (blk[0:nI].view(torch.int64).view(batch, k), blk[nI:2 * nI].view(torch.int64).view(batch, k),
                    blk[2 * nI:2 * nI + nD].view(torch.float32).view(batch, k),
                    blk[2 * nI + nD:].view(torch.float32).view(batch, k))

        self.I_img, self.I_txt, self.D_img, self.D_txt = views(self._res_host)
        Ii, It, Di, Dt = views(self._res_dev)
        self.perm = None if perm_img is None else perm_img.to(device=dev, dtype=torch.int32).contiguous()
        self.out: dict = {"I_img": Ii, "I_txt": It, "D_img": Di, "D_txt": Dt}
        self._host_out = {"I_img": self.I_img, "I_txt": self.I_txt, "D_img": self.D_img, "D_txt": self.D_txt}
        self._args = dict(topk=k, perm_img=self.perm, want_feats=want_feats, pool_mode=pool_mode, tau=tau)
        self.h2d_bytes = self.q_host.numel() * 4
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in (self.D_img, self.D_txt, self.I_img, self.I_txt))
        # warm-up queries: random directions, like real features. (All-zero queries tie every row of
        # the database, fail the exactness certificate and send the whole batch through the exact
        # fp32 fallback -- tens of milliseconds per warm-up run.)
        self.q_host.normal_(generator=torch.Generator().manual_seed(0))
        self.q_host.div_(self.q_host.norm(dim=1, keepdim=True))
        self.recaptures = -1
        self._capture()

    def _capture(self) -> None:
        if self._auto_layout:
            # time synchronised replays of both layouts (what run() does), alternating between them so
            # that clock drift hits both alike; keep the mapped layout unless the copy nodes are
            # clearly (> 2 %) faster on this box
            graphs = {}
            for cn in (False, True):
                self.copy_nodes = cn
                self._capture_one()
                graphs[cn] = self.graph
            st = torch.cuda.current_stream(self.q_dev.device)
            probe = {False: float("inf"), True: float("inf")}
            for rnd in range(5):
                for cn in (False, True):
                    g = graphs[cn]
                    g.replay()
                    st.synchronize()
                    t0 = time.perf_counter()
                    for _ in range(12):
                        g.replay()
                        st.synchronize()
                    probe[cn] = min(probe[cn], (time.perf_counter() - t0) / 12)
            self.layout_probe_us = {"host_io": probe[False] * 1e6, "copy_nodes": probe[True] * 1e6}
            self.copy_nodes = probe[True] < 0.98 * probe[False]
            del graphs
            self._auto_layout = False   # a re-capture (moved scratch) keeps the choice
            self.recaptures -= 2
        self._capture_one()

    def _capture_one(self) -> None:
        dev = self.q_dev.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):  # warm up: every scratch buffer reaches its final size before capture
                self._body()
            side.synchronize()
            self.ia.sync()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        # The graph holds raw addresses of the handles' scratch and row buffers. Any other call on
        # the same handles may move them (a larger batch or k, index.add): the handles count such
        # moves, and run() re-captures instead of replaying over freed memory.
        self._gen = (self.ia.generation, self.ib.generation)
        self.recaptures += 1

    def _body(self) -> None:
        if self.copy_nodes:
            self.q_dev.copy_(self.q_host, non_blocking=True)
            retrieve2(self.ia, self.ib, self.q_dev, out=self.out, **self._args)
            self._res_host.copy_(self._res_dev, non_blocking=True)
        else:
            retrieve2(self.ia, self.ib, self.q_host, out=self.out, host_out=self._host_out, **self._args)

    def run(self, q: Optional[torch.Tensor] = None, sync: bool = True) -> dict:
        if q is not None:
            self.q_host.copy_(q)
        if (self.ia.generation, self.ib.generation) != self._gen:
            self._capture()  # reads q_host only; the batch just written stays in place
        self.graph.replay()
        if sync:
            torch.cuda.current_stream(self.q_dev.device).synchronize()
        return self.out


class RetrievalPipeline:
    """Two `RetrievalStep`s over the same pair of indices, used alternately: `submit(q)` starts the
    step for one batch and returns at once, `wait(ticket)` hands back that batch's host results.
    The retrieval of batch i+1 does not depend on the results of batch i (the queries come from the
    encoder's forward of the new batch, src/trainer.py:52-58), so a loop may submit it before it
    consumes batch i: the graph launch and the completion wake-up of one step then hide under the
    GPU work of the other. Both graphs run on the caller's current stream, one after the other (the
    handles' scratch is shared: one search in flight per handle, as ever); queries and results are
    double-buffered, every step still reads its own queries from pinned host memory and returns
    its own (D, I) there."""

    def __init__(self, image_index: GpuIndexFlat, text_index: GpuIndexFlat, batch: int, **kw) -> None:
        self.steps = [RetrievalStep(image_index, text_index, batch, **kw) for _ in range(2)]
        self._events = [None, None]
        self._n = 0

    def submit(self, q: Optional[torch.Tensor] = None) -> int:
        slot = self._n & 1
        if self._events[slot] is not None:      # the slot's previous batch must be done with its buffers
            self._events[slot].synchronize()
        st = self.steps[slot]
        st.run(q, sync=False)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(st.q_dev.device))
        self._events[slot] = ev
        self._n += 1
        return self._n - 1

    def wait(self, ticket: int) -> "RetrievalStep":
        """Block until the batch submitted as `ticket` is complete; returns the step object whose
        D_img / I_img / D_txt / I_txt (pinned host) and `out` (device) hold its results. They stay
        valid until the second submit after this one."""
        slot = ticket & 1
        if ticket < self._n - 2 or ticket >= self._n:
            raise ValueError("ticket is not in flight (results are kept for the last two submits)")
        self._events[slot].synchronize()
        return self.steps[slot]


def get_extra_cap_features(feature, database, args=None, topk: int = 2):
    """src/trainer.py:262-283: top-`topk` text neighbours + their basenames (row-major order)."""
    image_base, text_base, basenames, image_index, text_index = (database[i] for i in range(5))
    if not _native(text_index):
        raise TypeError("database[4] must be a keds_b200 GPU index (no CPU path)")
    feature = feature / feature.norm(dim=1, keepdim=True)
    dev = torch.device("cuda", text_index.device)
    q = feature.detach().to(device=dev, dtype=torch.float32).contiguous()
    _, It = text_index.search(q, topk)
    feats = gather_rows(text_index, It, None)
    ids = It.cpu().numpy()
    names = [basenames[int(j)] for row in ids for j in row]
    return feats.to(feature.device), names


def retrieve_and_pool(feature: torch.Tensor, database, topk: int = 16, tau: float = 100.0):
    """Fused consumer primitive: search both databases, then softmax(tau * D)-weighted pool of the
    neighbours of each stream -> ([B, 1, d], [B, 1, d]).  The attn@v shape of
    src/model/model.py:69-73 with the weights supplied by the caller's scores."""
    image_index, text_index = database[3], database[4]
    dev = torch.device("cuda", image_index.device)
    q = feature.detach().to(device=dev, dtype=torch.float32).contiguous()
    o = retrieve2(image_index, text_index, q, topk, want_feats=False, pool_mode=POOL_SOFTMAX, tau=tau)
    return o["pool_img"].unsqueeze(1), o["pool_txt"].unsqueeze(1)
