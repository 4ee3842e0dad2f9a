"""GPU (-m gpu): the gathered contrastive loss through the C ABI against the
float64 oracle (oracle/clip_loss_oracle.py, pinned to torch autograd of the reference statements).

Tolerance: products run on tf32 tensor cores with a hi/lo operand split (error ~2^-22 relative),
everything else in fp32; asserted: loss within 2e-5 relative, gradients within 2e-5 of their
largest entry, the scale gradient (a cancelling sum of N^2 fp32 terms) within 2e-4 relative.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from keds_b200 import contrastive as kc  # noqa: E402
from oracle import clip_loss_oracle as lorc  # noqa: E402

TOL = 2e-5


def feats(n, d, seed, noise=6.0):
    rng = np.random.default_rng(seed)
    I = rng.standard_normal((n, d))
    T = I + noise * rng.standard_normal((n, d))      # pairs stay the best match, but not by a mile
    I /= np.linalg.norm(I, axis=1, keepdims=True)
    T /= np.linalg.norm(T, axis=1, keepdims=True)
    return I.astype(np.float32), T.astype(np.float32)


def close(got, want, tol=TOL):
    return np.abs(np.asarray(got, dtype=np.float64) - want).max() <= tol * max(1e-30, np.abs(want).max())


@pytest.mark.parametrize("N,d,row0,B,scale", [(1024, 768, 256, 128, 100.0), (128, 768, 0, 128, 14.2857),
                                               (260, 64, 100, 60, 50.0), (8, 16, 4, 4, 3.0)])
def test_loss_and_gradients_match_oracle(N, d, row0, B, scale):
    I, T = feats(N, d, seed=N + d)
    loss, dI, dT, ds = kc.clip_loss_forward_backward(torch.from_numpy(I).cuda(), torch.from_numpy(T).cuda(),
                                                     torch.tensor(scale), row0, B)
    kc.check(0)
    wl, wdI, wdT, wds = lorc.clip_loss(I, T, scale, row0, B)
    assert wl > 1e-3 and abs(float(loss) - wl) <= TOL * abs(wl), (float(loss), wl)
    assert close(dI.cpu().numpy(), wdI) and close(dT.cpu().numpy(), wdT)
    # d loss / d scale = sum G * L: ~N^2 terms of both signs accumulated in fp32
    assert abs(float(ds) - wds) <= 2e-4 * abs(wds) + 1e-6, (float(ds), wds)


def test_reference_statements_golden(golden_dir):
    # two simulated ranks, the reference's own statements through torch autograd (float64)
    g = np.load(os.path.join(golden_dir, "clip_loss.npz"))
    world, scale = int(g["world"]), float(g["scale"])
    I_all = np.concatenate([g[f"I{r}"] for r in range(world)]).astype(np.float32)
    T_all = np.concatenate([g[f"T{r}"] for r in range(world)]).astype(np.float32)
    B = g["I0"].shape[0]
    for r in range(world):
        loss, dI, dT, ds = kc.clip_loss_forward_backward(torch.from_numpy(I_all).cuda(), torch.from_numpy(T_all).cuda(),
                                                         torch.tensor(scale), r * B, B)
        assert abs(float(loss) - float(g[f"loss{r}"])) < 1e-5
        assert close(dI.cpu().numpy(), g[f"dI{r}"], 1e-4) and close(dT.cpu().numpy(), g[f"dT{r}"], 1e-4)
        assert abs(float(ds) - float(g[f"dscale{r}"])) < 1e-5


def test_autograd_function_single_process():
    # no process group: N = B; gradients flow to both feature tensors and to logit_scale
    I, T = feats(128, 768, seed=5)
    Ii = torch.from_numpy(I).cuda().requires_grad_(True)
    Tt = torch.from_numpy(T).cuda().requires_grad_(True)
    log_scale = torch.tensor(np.log(100.0), device="cuda", requires_grad=True)
    loss = kc.gathered_clip_loss(Ii, Tt, log_scale.exp()) * 3.0
    loss.backward()
    wl, wdI, wdT, wds = lorc.clip_loss(I, T, 100.0)
    assert abs(float(loss.detach()) - 3 * wl) <= 3 * TOL * abs(wl)
    assert close(Ii.grad.cpu().numpy(), 3 * wdI) and close(Tt.grad.cpu().numpy(), 3 * wdT)
    assert abs(float(log_scale.grad) - 3 * wds * 100.0) <= 1e-4 * abs(3 * wds * 100.0) + 1e-6
    # against torch's own autograd of the reference statements on the GPU (fp32)
    Ir = torch.from_numpy(I).cuda().requires_grad_(True)
    Tr = torch.from_numpy(T).cuda().requires_grad_(True)
    torch.backends.cuda.matmul.allow_tf32 = False
    logits = 100.0 * Ir @ Tr.t()
    gt = torch.arange(128, device="cuda")
    ref = (torch.nn.functional.cross_entropy(logits, gt) + torch.nn.functional.cross_entropy(logits.t(), gt)) / 2
    (3.0 * ref).backward()
    assert abs(float(loss.detach()) - 3 * float(ref.detach())) < 1e-4
    assert (Ii.grad - Ir.grad).abs().max() < 1e-4 * Ir.grad.abs().max() + 1e-7


def test_clip_loss_argument_errors():
    I = torch.zeros(6, 16, device="cuda")
    with pytest.raises(RuntimeError):
        kc.clip_loss_forward_backward(I, I, torch.tensor(1.0), 0, 6)          # N not a multiple of 4
    with pytest.raises(TypeError):
        kc.clip_loss_forward_backward(I.cpu(), I.cpu(), torch.tensor(1.0), 0, 6)
    J = torch.zeros(8, 16, device="cuda")
    with pytest.raises(RuntimeError):
        kc.clip_loss_forward_backward(J, J, torch.tensor(1.0), 4, 8)          # local rows outside [0, N)
