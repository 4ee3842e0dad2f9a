"""The contrastive loss over the gathered features, on the native path (forward and backward).

The reference (src/trainer.py:85-135,164), per training step and rank:

    dist.all_gather(gathered_image_features, image_features)       # 2 collectives
    dist.all_gather(gathered_text_features, text_features)
    all_image_features = cat([image_features] + others)             # local rows first
    logits_per_image = logit_scale * all_image_features @ all_text_features.t()
    total_loss = (CE(logits_per_image, arange) + CE(logits_per_image.t(), arange)) / 2

`gathered_clip_loss` returns the same scalar with the same gradients (to the LOCAL image / text
features and to logit_scale -- the gathered copies carry no gradient in the reference either).
One all-gather of the packed [image | text] features, then one native call that produces the loss
and all three gradients (`keds_clip_loss_forward_backward`); autograd's backward only scales them.
The row order differs from the reference (rank order instead of local-first): the loss pairs row i
with column i, so any order applied to both sides gives the same value.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _capi
from .index import _stream_ptr

_handles = {}


def _handle(device: int):
    h = _handles.get(device)
    if h is None:
        lib = _capi.load()
        h = C.c_void_p()
        _capi.check(lib.keds_clip_loss_create(int(device), C.byref(h)))
        _handles[device] = h
    return h


def gather_features(image_features: torch.Tensor, text_features: torch.Tensor, group=None
                    ) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """All ranks' features in rank order and the first row of this rank: one collective on the
    packed [B, 2, d] block instead of the reference's two (src/trainer.py:100-101). Without an
    initialised process group (single process) the inputs come back unchanged."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return image_features, text_features, 0
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B, d = image_features.shape
    packed = torch.stack([image_features.detach(), text_features.detach()], dim=1).contiguous()   # [B, 2, d]
    out = torch.empty((world * B, 2, d), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    return out[:, 0, :].contiguous(), out[:, 1, :].contiguous(), rank * B


def clip_loss_forward_backward(I_all: torch.Tensor, T_all: torch.Tensor, logit_scale: torch.Tensor, row0: int,
                               n_local: int, want_grad: bool = True):
    """(loss, dI_local, dT_local, dscale) as device tensors (gradients None if not wanted)."""
    lib = _capi.load()
    for t in (I_all, T_all):
        if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2):
            raise TypeError("features must be 2-D CUDA float32 tensors (there is no CPU path)")
    if I_all.shape != T_all.shape:
        raise ValueError("image and text features must have the same shape")
    I_all, T_all = I_all.contiguous(), T_all.contiguous()
    N, d = I_all.shape
    dev = I_all.device
    scale = logit_scale.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()  # stays on the device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    dscale = torch.empty((), dtype=torch.float32, device=dev)
    dI = torch.empty((n_local, d), dtype=torch.float32, device=dev) if want_grad else None
    dT = torch.empty((n_local, d), dtype=torch.float32, device=dev) if want_grad else None
    _capi.check(lib.keds_clip_loss_forward_backward(
        _handle(dev.index), I_all.data_ptr(), T_all.data_ptr(), N, d, int(row0), int(n_local), scale.data_ptr(),
        loss.data_ptr(), dI.data_ptr() if want_grad else 0, dT.data_ptr() if want_grad else 0, dscale.data_ptr(),
        _stream_ptr(dev.index)))
    return loss, dI, dT, dscale


class _GatheredClipLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image_features, text_features, logit_scale, group):
        I_all, T_all, row0 = gather_features(image_features, text_features, group)
        need = any(ctx.needs_input_grad[:3])
        loss, dI, dT, dscale = clip_loss_forward_backward(I_all, T_all, logit_scale, row0,
                                                          image_features.shape[0], want_grad=need)
        if need:
            ctx.save_for_backward(dI, dT, dscale)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        dI, dT, dscale = ctx.saved_tensors
        gi = grad_out * dI if ctx.needs_input_grad[0] else None
        gt = grad_out * dT if ctx.needs_input_grad[1] else None
        gs = grad_out * dscale if ctx.needs_input_grad[2] else None
        return gi, gt, gs, None


def gathered_clip_loss(image_features: torch.Tensor, text_features: torch.Tensor, logit_scale: torch.Tensor,
                       group=None) -> torch.Tensor:
    """total_loss of src/trainer.py:164 for this rank's [B, d] image / text features (CUDA float32,
    already normalised as in :80-81) and the scalar tensor logit_scale (= model.logit_scale.exp().mean(),
    :86-87). Differentiable in all three."""
    if not torch.is_tensor(logit_scale):
        logit_scale = torch.tensor(float(logit_scale), device=image_features.device)
    return _GatheredClipLoss.apply(image_features, text_features, logit_scale, group)


def check(device: int = 0) -> None:
    """Synchronise and raise if a kernel of the loss reported a pipeline error."""
    _capi.check(_capi.load().keds_clip_loss_check(_handle(device), _stream_ptr(device)))
