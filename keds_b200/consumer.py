"""Forward pass of the modules that consume the retrieved neighbours, on the native path.

The reference evaluates, per batch (src/trainer.py:59-69; src/eval_utils.py:378-383, 515-519,
661-668, 806-810, 943-947):

    mapped   = img2text(image_features)                       # IM2TEXT, src/model/model.py:104-123
    nb_img   = img2text(topk_image_features)                  # [B, k, 768]
    nb_txt   = img2text(topk_text_features)
    fused    = retrieval_fuse(mapped.unsqueeze(1), nb_img, nb_img)     # CrossFormer, :81-101
    text_c   = text_condition(mapped.unsqueeze(1), nb_txt, nb_txt)
    tokens   = torch.cat([fused, text_c, mapped.unsqueeze(1)], dim=1)  # [B, 3, 768]

as ~60 small PyTorch kernels on `[B, k, 768]` tensors that first travelled GPU -> CPU -> GPU.
`NeighbourConsumer` takes the three modules' weights once and produces `tokens` from the query
features and the neighbour *ids* (the rows are gathered from the resident databases).  Forward
only: training, where the modules learn, keeps the PyTorch modules.

Weights are taken from `state_dict()`s with the reference's own parameter names, so
`NeighbourConsumer.from_modules(img2text, retrieval_fuse, text_condition)` works on the reference's
module instances unchanged.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Optional

import torch

from . import _capi
from .index import GpuIndexFlat, _stream_ptr

KIND_MLP, KIND_TO_Q, KIND_TO_K, KIND_TO_V, KIND_TO_OUT = 0, 1, 2, 3, 4
STACK_IMAGE, STACK_TEXT = 0, 1


def _count_prefix(sd: Mapping[str, torch.Tensor], fmt: str) -> int:
    n = 0
    while fmt.format(n) in sd:
        n += 1
    return n


class NeighbourConsumer:
    """img2text + retrieval_fuse + text_condition in one launch sequence (tf32 tensor cores)."""

    def __init__(self, img2text_sd: Mapping[str, torch.Tensor], retrieval_fuse_sd: Mapping[str, torch.Tensor],
                 text_condition_sd: Mapping[str, torch.Tensor], heads: int = 8, device: int = 0) -> None:
        self._lib = _capi.load()
        self._h = C.c_void_p()
        self.device = int(device)
        # IM2TEXT: layers.{i}.0 = Linear (+ Dropout + ReLU), fc_out (src/model/model.py:107-118)
        n_hidden = _count_prefix(img2text_sd, "layers.{}.0.weight")
        if n_hidden < 1 or "fc_out.weight" not in img2text_sd:
            raise ValueError("img2text state_dict: expected layers.<i>.0.weight and fc_out.weight")
        w0 = img2text_sd["layers.0.0.weight"]
        d_mid, d_in = int(w0.shape[0]), int(w0.shape[1])
        d_tok = int(img2text_sd["fc_out.weight"].shape[0])
        # CrossFormer: cross_layers.{l}.to_q / to_k / to_v / to_out.0 (src/model/model.py:46-55, 93)
        n_layers = _count_prefix(retrieval_fuse_sd, "cross_layers.{}.to_q.weight")
        if n_layers < 1 or _count_prefix(text_condition_sd, "cross_layers.{}.to_q.weight") != n_layers:
            raise ValueError("retrieval_fuse / text_condition state_dicts: cross_layers.<l>.to_q.weight missing "
                             "or different depths")
        inner = int(retrieval_fuse_sd["cross_layers.0.to_q.weight"].shape[0])
        if inner % heads:
            raise ValueError(f"inner width {inner} is not a multiple of heads={heads}")
        if "cross_layers.0.to_out.0.weight" not in retrieval_fuse_sd:
            raise ValueError("CrossAttention without output projection (heads == 1 and dim_head == q_dim) "
                             "is not supported")
        self.d_in, self.d_mid, self.d_tok = d_in, d_mid, d_tok
        self.n_hidden, self.n_layers, self.heads, self.dim_head = n_hidden, n_layers, heads, inner // heads
        _capi.check(self._lib.keds_consumer_create(d_in, d_mid, d_tok, n_hidden, n_layers, heads, inner // heads,
                                                   self.device, C.byref(self._h)))
        for i in range(n_hidden):
            self._set(KIND_MLP, 0, i, img2text_sd[f"layers.{i}.0.weight"], img2text_sd.get(f"layers.{i}.0.bias"))
        self._set(KIND_MLP, 0, n_hidden, img2text_sd["fc_out.weight"], img2text_sd.get("fc_out.bias"))
        for stack, sd in ((STACK_IMAGE, retrieval_fuse_sd), (STACK_TEXT, text_condition_sd)):
            for l in range(n_layers):
                p = f"cross_layers.{l}."
                self._set(KIND_TO_Q, stack, l, sd[p + "to_q.weight"], sd.get(p + "to_q.bias"))
                self._set(KIND_TO_K, stack, l, sd[p + "to_k.weight"], sd.get(p + "to_k.bias"))
                self._set(KIND_TO_V, stack, l, sd[p + "to_v.weight"], sd.get(p + "to_v.bias"))
                self._set(KIND_TO_OUT, stack, l, sd[p + "to_out.0.weight"], sd.get(p + "to_out.0.bias"))
        _capi.check(self._lib.keds_consumer_finalize(self._h))

    @classmethod
    def from_modules(cls, img2text, retrieval_fuse, text_condition, device: int = 0) -> "NeighbourConsumer":
        """Build from the reference's module instances (IM2TEXT, CrossFormer, CrossFormer)."""
        heads = int(retrieval_fuse.cross_layers[0].heads)
        return cls(img2text.state_dict(), retrieval_fuse.state_dict(), text_condition.state_dict(),
                   heads=heads, device=device)

    def _set(self, kind: int, stack: int, layer: int, W: torch.Tensor, b: Optional[torch.Tensor]) -> None:
        W = W.detach().to(dtype=torch.float32).contiguous()
        bp = 0
        if b is not None:
            b = b.detach().to(dtype=torch.float32).contiguous()
            bp = b.data_ptr()
        if W.is_cuda:  # the native copy runs on the legacy stream: finish whatever produced W / b first
            torch.cuda.current_stream(W.device).synchronize()
        _capi.check(self._lib.keds_consumer_set_linear(self._h, kind, stack, layer, W.data_ptr(), bp,
                                                       int(W.shape[0]), int(W.shape[1])))

    def __call__(self, feature: torch.Tensor, image_index: GpuIndexFlat, text_index: GpuIndexFlat,
                 I_img: torch.Tensor, I_txt: torch.Tensor, perm: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """tokens [B, 3, d_tok] = cat(fused, text_conditioned, mapped) for query features [B, d_in]
        and neighbour ids I_img / I_txt [B, k] (int64, device) into the two resident databases."""
        if not (feature.is_cuda and feature.dtype == torch.float32 and feature.dim() == 2
                and feature.shape[1] == self.d_in):
            raise TypeError(f"feature must be a CUDA float32 [B, {self.d_in}] tensor")
        for I in (I_img, I_txt):
            if not (I.is_cuda and I.dtype == torch.int64 and I.dim() == 2):
                raise TypeError("neighbour ids must be CUDA int64 [B, k] tensors")
        if I_img.shape != I_txt.shape or I_img.shape[0] != feature.shape[0]:
            raise ValueError("feature, I_img and I_txt disagree on B or k")
        if image_index.d != self.d_in or text_index.d != self.d_in:
            raise ValueError("database width differs from the consumer's input width")
        feature, I_img, I_txt = feature.contiguous(), I_img.contiguous(), I_txt.contiguous()
        B, k = I_img.shape
        if out is None:
            out = torch.empty((B, 3, self.d_tok), dtype=torch.float32, device=feature.device)
        pp = 0
        if perm is not None:
            perm = perm.to(device=feature.device, dtype=torch.int32).contiguous()
            assert perm.numel() == k
            pp = perm.data_ptr()
        _capi.check(self._lib.keds_consumer_forward(
            self._h, feature.data_ptr(), image_index.rows_ptr(), image_index.ntotal, text_index.rows_ptr(),
            text_index.ntotal, I_img.data_ptr(), I_txt.data_ptr(), pp, B, k, out.data_ptr(),
            _stream_ptr(self.device)))
        return out

    def from_features(self, image_features: torch.Tensor, database, topk: int = 16,
                      shuffle: bool = True) -> torch.Tensor:
        """The whole block of src/trainer.py:53-69 / src/eval_utils.py:373-383 for one batch:
        retrieve the topk image / text neighbours of `image_features` from `database`
        (KnowledgeBase or the reference's 5-item sequence with native indices) and return
        tokens [B, 3, d_tok]. The search normalises its copy of the features (src/trainer.py:206);
        img2text sees them as given. shuffle: the batch-shared randperm of :218-219 (it cannot
        change the result beyond fp32 summation order)."""
        from .index import search2

        image_index, text_index = database[3], database[4]
        feats = image_features.detach().to(device=torch.device("cuda", self.device), dtype=torch.float32).contiguous()
        q = feats / feats.norm(dim=1, keepdim=True)
        (_, I_img), (_, I_txt) = search2(image_index, text_index, q, topk)
        perm = torch.randperm(topk) if shuffle else None
        return self(feats, image_index, text_index, I_img, I_txt, perm)

    def check(self) -> int:
        """Synchronise and raise if a kernel reported a pipeline error; returns kernels launched so far."""
        n = C.c_int64(0)
        _capi.check(self._lib.keds_consumer_check(self._h, _stream_ptr(self.device), C.byref(n)))
        return int(n.value)

    def set_debug(self, on: bool) -> None:
        _capi.check(self._lib.keds_consumer_set_debug(self._h, int(bool(on))))

    def debug_timeline(self, launch: int, n_ctas: int):
        """[n_ctas, 5] uint64 ns timestamps {start, prologue, dependency, accumulator, end} of the
        `launch`-th k_linear_tf32 launch of the last forward (debug on)."""
        import numpy as np

        out = np.zeros((n_ctas, 5), dtype=np.uint64)
        _capi.check(self._lib.keds_consumer_debug_timeline(self._h, launch, out.ctypes.data, n_ctas))
        return out

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.keds_consumer_free(self._h)
            self._h = C.c_void_p()

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
