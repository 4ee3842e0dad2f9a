"""CPU oracle for the gathered contrastive loss.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py may import this module; the product
(keds_b200/) never does and has no CPU path.

numpy float64 restatement of src/trainer.py:85-135,164 (aggregate branch):
    logits_per_image = logit_scale * all_image_features @ all_text_features.t()        (:126)
    total_loss = (CE(logits_per_image, arange) + CE(logits_per_image.t(), arange)) / 2  (:127-129,164)
with CE = nn.CrossEntropyLoss() (mean reduction), and its gradients with respect to this rank's
rows and to logit_scale (the gathered copies carry no gradient: dist.all_gather is not
differentiable and only `image_features` / `text_features` themselves enter the cat at :103-112).

PARITY PINNING: tests/test_oracle_golden.py checks it against tests/golden/clip_loss.npz, produced by
oracle/make_golden_clip_loss.py with the reference's own statements run through torch
(nn.CrossEntropyLoss + autograd, float64) for two simulated ranks, local-rows-first as at :103-112.
"""
from __future__ import annotations

import numpy as np


def clip_loss(I_all, T_all, scale: float, row0: int = 0, n_local: int | None = None):
    """(loss, dI_local, dT_local, dscale) for features in any common row order."""
    I = np.asarray(I_all, dtype=np.float64)
    T = np.asarray(T_all, dtype=np.float64)
    N = I.shape[0]
    n_local = N if n_local is None else n_local
    L = I @ T.T
    P = scale * L
    mr = P.max(axis=1, keepdims=True)
    lse_r = (mr + np.log(np.exp(P - mr).sum(axis=1, keepdims=True)))[:, 0]
    mc = P.max(axis=0, keepdims=True)
    lse_c = (mc + np.log(np.exp(P - mc).sum(axis=0, keepdims=True)))[0]
    diag = np.diag(P)
    loss = ((lse_r - diag).mean() + (lse_c - diag).mean()) / 2.0
    G = (np.exp(P - lse_r[:, None]) + np.exp(P - lse_c[None, :]) - 2.0 * np.eye(N)) / (2.0 * N)   # dloss/dP
    sl = slice(row0, row0 + n_local)
    dI = scale * G[sl, :] @ T
    dT = scale * G[:, sl].T @ I
    dscale = float((G * L).sum())
    return float(loss), dI, dT, dscale
