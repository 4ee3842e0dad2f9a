"""GPU (-m gpu): the neighbour consumer in TRAINING mode (src/trainer.py:59-69 with
the backward of :462-474) through the C ABI -- tokens and the gradient of every parameter -- against
the float64 oracle (oracle/consumer_train_oracle.py, pinned to the reference's own modules by
tests/golden/consumer_train.npz).

Tolerance: every product runs on tf32 tensor cores (10-bit mantissa operands, fp32 accumulation);
the reference trains in fp32 (fp16 under amp autocast, src/trainer.py:462-465). Asserted per
tensor: max |got - oracle| <= 1e-2 * max |oracle| (measured: 5e-4 at the reference's widths on
unit-norm inputs, 5e-3 on unnormalised inputs after an SGD step), with the ReLU gates of the native
forward (see check_against_oracle).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from keds_b200 import faiss_compat as faiss  # noqa: E402
from keds_b200.consumer import NeighbourConsumer, TrainableNeighbourConsumer  # noqa: E402
from keds_b200.index import GpuIndexFlat  # noqa: E402
from oracle import consumer_oracle as corc  # noqa: E402
from oracle import consumer_train_oracle as cto  # noqa: E402

TOK_TOL, GRAD_TOL = 4e-3, 1e-2


def to_torch(sd):
    return {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)) for k, v in sd.items()}


def index_of(rows):
    ix = GpuIndexFlat(rows.shape[1], faiss.METRIC_INNER_PRODUCT, 0)
    ix.add(np.ascontiguousarray(rows, dtype=np.float32))
    return ix


def rel(got, want, floor=1e-30):
    """max |got - want| relative to the tensor's own scale -- but never below `floor`: the to_k
    biases have an exactly zero gradient (a bias on every key shifts all scores of a softmax row
    alike), so their error is measured against the scale of the other gradients."""
    return float(np.abs(got - want).max() / max(np.abs(want).max(), floor))


def run(sds, heads, feat, base_img, base_txt, I_img, I_txt, dtok, masks=None, dropout=0.0):
    mod = TrainableNeighbourConsumer(to_torch(sds[0]), to_torch(sds[1]), to_torch(sds[2]), heads=heads, device=0,
                                     dropout=dropout)
    ix_i, ix_t = index_of(base_img), index_of(base_txt)
    f = torch.from_numpy(feat.astype(np.float32)).cuda()
    Ii, It = torch.from_numpy(I_img).cuda(), torch.from_numpy(I_txt).cuda()
    return mod, ix_i, ix_t, f, Ii, It


def check_against_oracle(sds, heads, feat, base_img, base_txt, I_img, I_txt, dtok, masks=None):
    mod, ix_i, ix_t, f, Ii, It = run(sds, heads, feat, base_img, base_txt, I_img, I_txt, dtok)
    from keds_b200.consumer import _ConsumerFn
    mk = None if masks is None else [None if m is None else torch.from_numpy(m.astype(np.float32)).cuda() for m in masks]
    tokens = _ConsumerFn.apply(mod, f, ix_i, ix_t, Ii, It, None, mk, mod.flat)
    tokens.backward(torch.from_numpy(dtok.astype(np.float32)).cuda())
    assert mod.check() > 0
    B, k = I_img.shape
    nb_i = base_img[I_img.reshape(-1)].reshape(B, k, -1)
    nb_t = base_txt[I_txt.reshape(-1)].reshape(B, k, -1)
    # The ReLU gates come from the native forward: a tf32 forward and the float64 oracle disagree on
    # the few units whose pre-activation is within rounding of zero (measured: ~4e-4 of the first
    # layer's units at full width), and one flipped gate per column already moves a bias gradient by
    # several per cent. They must agree almost everywhere, though.
    M = B * (1 + 2 * k)
    n_hidden = sum(1 for name in sds[0] if name.endswith(".0.weight"))
    gates = [(mod.debug_hidden(i, M) > 0).double().cpu().numpy() for i in range(n_hidden)]
    _, free = cto.tokens_and_grads(sds[0], sds[1], sds[2], heads, feat, nb_i, nb_t, dtok, masks)
    want_tok, want = cto.tokens_and_grads(sds[0], sds[1], sds[2], heads, feat, nb_i, nb_t, dtok, masks, gates)
    assert rel(want["img2text/fc_out.weight"], free["img2text/fc_out.weight"]) < 1e-2   # same network, same gates almost everywhere
    errs = {"tokens": rel(tokens.detach().cpu().numpy(), want_tok)}
    assert errs["tokens"] < TOK_TOL, errs
    gi, gf, gc = mod.state_dicts(grads=True)
    scale = 1e-3 * max(float(np.abs(v).max()) for v in want.values())
    for prefix, gsd in (("img2text", gi), ("retrieval_fuse", gf), ("text_condition", gc)):
        for name, t in gsd.items():
            errs[f"{prefix}/{name}"] = rel(t.cpu().numpy(), want[f"{prefix}/{name}"], scale)
    worst = max(errs, key=errs.get)
    assert errs[worst] < GRAD_TOL, "\n".join(f"{v:.3e}  {k_}" for k_, v in sorted(errs.items(), key=lambda kv: -kv[1])[:14])
    return errs, mod


def test_reference_module_gradients_small_widths(golden_dir):
    """against what the reference's own modules produced in train mode (ragged widths 48 / 32 / 40)"""
    g = np.load(os.path.join(golden_dir, "consumer_train.npz"))
    heads = int(g["dims"][5])
    sds = [{k.split("/", 1)[1]: g[k] for k in g.files if k.startswith(p + "/")}
           for p in ("img2text", "retrieval_fuse", "text_condition")]
    B, k, d_in = g["topk_image"].shape
    base_img = g["topk_image"].reshape(B * k, d_in).astype(np.float32)
    base_txt = g["topk_text"].reshape(B * k, d_in).astype(np.float32)
    I = np.arange(B * k, dtype=np.int64).reshape(B, k)
    errs, mod = check_against_oracle(sds, heads, g["feat"].astype(np.float32), base_img, base_txt, I, I, g["dtokens"])
    gi, gf, gc = mod.state_dicts(grads=True)
    scale = 1e-3 * max(float(np.abs(g[k_]).max()) for k_ in g.files if k_.startswith("grad/"))
    for prefix, gsd in (("img2text", gi), ("retrieval_fuse", gf), ("text_condition", gc)):
        for name, t in gsd.items():
            assert rel(t.cpu().numpy(), g[f"grad/{prefix}/{name}"], scale) < GRAD_TOL, (prefix, name)


@pytest.mark.parametrize("B,k,with_masks", [(128, 16, True), (37, 16, False), (130, 5, True)])
def test_full_width_gradients_match_oracle(B, k, with_masks):
    """the reference's sizes (768 -> 512 -> 512 -> 768 MLP, 3 layers of 8 heads x 64) with dropout
    masks on both hidden layers"""
    sds = corc.random_state_dicts(768, 512, 768, 2, 3, 8, 64, seed=11 + B)
    rng = np.random.default_rng(B * 7 + k)
    n = 3000
    base_img = rng.standard_normal((n, 768)).astype(np.float32)
    base_img /= np.linalg.norm(base_img, axis=1, keepdims=True)
    base_txt = rng.standard_normal((n, 768)).astype(np.float32)
    base_txt /= np.linalg.norm(base_txt, axis=1, keepdims=True)
    feat = rng.standard_normal((B, 768)).astype(np.float32)
    feat /= np.linalg.norm(feat, axis=1, keepdims=True)
    I_img = rng.integers(0, n, (B, k)).astype(np.int64)
    I_txt = rng.integers(0, n, (B, k)).astype(np.int64)
    dtok = rng.standard_normal((B, 3, 768)).astype(np.float32)
    masks = None
    if with_masks:
        M = B * (1 + 2 * k)
        masks = [((rng.random((M, 512)) < 0.9) / 0.9).astype(np.float32) for _ in range(2)]
    errs, _ = check_against_oracle(sds, 8, feat, base_img, base_txt, I_img, I_txt, dtok, masks)
    print("worst relative errors:", sorted(errs.items(), key=lambda kv: -kv[1])[:4])


def test_optimizer_updates_are_read_through_and_eval_matches_the_forward_only_consumer():
    """the flat Parameter IS the weight storage: an optimiser step changes the next forward without
    any upload, the transposed copies used by dX are refreshed, and eval() reproduces the
    forward-only consumer bit for bit on the same weights"""
    sds = corc.random_state_dicts(256, 128, 256, 2, 2, 4, 32, seed=3)
    rng = np.random.default_rng(9)
    n, B, k = 500, 40, 8
    base = rng.standard_normal((n, 256)).astype(np.float32)
    feat = rng.standard_normal((B, 256)).astype(np.float32)
    I1 = rng.integers(0, n, (B, k)).astype(np.int64)
    I2 = rng.integers(0, n, (B, k)).astype(np.int64)
    dtok = rng.standard_normal((B, 3, 256)).astype(np.float32)
    mod, ix_i, ix_t, f, Ii, It = run(sds, 4, feat, base, base, I1, I2, dtok)
    opt = torch.optim.SGD(mod.parameters(), lr=0.05)
    mod.eval()
    t0 = mod(f, ix_i, ix_t, Ii, It)
    ref = NeighbourConsumer(*mod.state_dicts(), heads=4, device=0)
    assert torch.equal(t0.detach(), ref(f, ix_i, ix_t, Ii, It))
    (t0 * torch.from_numpy(dtok).cuda()).sum().backward()
    opt.step()
    opt.zero_grad()
    # after the step: compare forward and backward with the oracle on the UPDATED weights
    new_sds = [{k_: v.cpu().numpy() for k_, v in sd.items()} for sd in mod.state_dicts()]
    t1 = mod(f, ix_i, ix_t, Ii, It)
    t1.backward(torch.from_numpy(dtok).cuda())
    nb_i = base[I1.reshape(-1)].reshape(B, k, -1)
    nb_t = base[I2.reshape(-1)].reshape(B, k, -1)
    gates = [(mod.debug_hidden(i, B * (1 + 2 * k)) > 0).double().cpu().numpy() for i in range(2)]
    want_tok, want = cto.tokens_and_grads(new_sds[0], new_sds[1], new_sds[2], 4, feat, nb_i, nb_t, dtok, None, gates)
    assert rel(t1.detach().cpu().numpy(), want_tok) < TOK_TOL
    assert not torch.equal(t1.detach(), t0.detach())
    gi, gf, gc = mod.state_dicts(grads=True)
    scale = 1e-3 * max(float(np.abs(v).max()) for v in want.values())
    for prefix, gsd in (("img2text", gi), ("retrieval_fuse", gf), ("text_condition", gc)):
        for name, t in gsd.items():
            assert rel(t.cpu().numpy(), want[f"{prefix}/{name}"], scale) < GRAD_TOL, (prefix, name)
    # train() mode draws dropout masks: outputs differ from eval, gradients still flow
    mod.train()
    mod.dropout = 0.5
    t2 = mod(f, ix_i, ix_t, Ii, It)
    assert not torch.equal(t2.detach(), t1.detach())
    t2.sum().backward()
    assert torch.isfinite(mod.flat.grad).all() and float(mod.flat.grad.abs().max()) > 0
