"""CPU oracle for the KEDs knowledge-retrieval hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (keds_b200/) never does and has no CPU path.

What it restates, with the reference lines it follows (paths relative to the KEDs tree):

  * flat exact search  -- the arithmetic lives in Faiss (third party, NOT in the tree, not
    installable here; the only pin anywhere is faiss-gpu=1.4.0 in
    src/third_party/open_clip/environment.yml:27).  Published algorithm of IndexFlatIP /
    IndexFlatL2: brute-force inner product / squared Euclidean distance of every query against
    every row, k best per query, best first, rows past ntotal padded with label -1 and
    -FLT_MAX (IP) / +FLT_MAX (L2).  Call sites: src/trainer.py:213,221,271,
    src/eval_utils.py:169,177; index construction src/main.py:72-83.  The reference's own torch
    restatement of the same search is src/trainer.py:246-257 (matmul -> topk -> gather).
  * neighbour gather + batch-shared randperm shuffle   src/trainer.py:214-230
  * weighted pool (attn @ v shape)                      src/model/model.py:69-73
  * gallery metrics                                     src/eval_utils.py:1008-1134

PARITY PINNING.  The reference has no tests, fixtures or golden vectors, and Faiss cannot be run
here, so the Faiss boundary itself is "parity unpinned" by the reference.  What IS pinned: this
oracle is checked (tests/test_oracle_golden.py) against outputs of the reference's own functions
executed unchanged from /root/reference via AST extraction (oracle/make_golden.py ->
tests/golden/*.npz): get_retrieved_features(use_faiss=False) and both branches with a stand-in
index, and get_metrics_coco / _fashion / _cirr / _imgnet.

Ties are unspecified upstream (Faiss heap order, unstable torch.argsort); the oracle fixes
(score descending, label ascending) and comparisons allow set equality inside near-tie groups.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

FLT_MAX = np.finfo(np.float32).max


def _topk_desc(scores: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Top-k per row by (score desc, index asc). scores: [B, N] float64."""
    B, N = scores.shape
    kk = min(k, N)
    if kk == 0:
        return np.empty((B, 0)), np.empty((B, 0), np.int64)
    if kk < N:
        # everything >= the kk-th value is a contender; resolve ties by index on that small set
        part = np.partition(scores, N - kk, axis=1)[:, N - kk]
    else:
        part = np.full(B, -np.inf)
    I = np.empty((B, kk), np.int64)
    D = np.empty((B, kk), scores.dtype)
    for b in range(B):
        cand = np.nonzero(scores[b] >= part[b])[0]
        order = np.lexsort((cand, -scores[b, cand]))[:kk]
        I[b] = cand[order]
        D[b] = scores[b, I[b]]
    return D, I


def search(db: np.ndarray, q: np.ndarray, k: int, metric: str = "ip",
           chunk: int = 1024) -> Tuple[np.ndarray, np.ndarray]:
    """Exact flat search in float64. Returns (D float32 [B,k], I int64 [B,k]), Faiss conventions."""
    db64 = np.asarray(db, np.float64)
    q64 = np.asarray(q, np.float64)
    B, N = q64.shape[0], db64.shape[0]
    D = np.full((B, k), -FLT_MAX if metric == "ip" else FLT_MAX, np.float32)
    I = np.full((B, k), -1, np.int64)
    if N == 0 or B == 0:
        return D, I
    xn = (db64 * db64).sum(1) if metric == "l2" else None
    for lo in range(0, B, chunk):
        qq = q64[lo:lo + chunk]
        s = qq @ db64.T
        if metric == "l2":
            dist = (qq * qq).sum(1)[:, None] + xn[None, :] - 2.0 * s
            dist = np.maximum(dist, 0.0)
            d, i = _topk_desc(-dist, k)
            d = -d
        else:
            d, i = _topk_desc(s, k)
        kk = i.shape[1]
        D[lo:lo + chunk, :kk] = d.astype(np.float32)
        I[lo:lo + chunk, :kk] = i
    return D, I


def search_f32_blas(db: np.ndarray, q: np.ndarray, k: int, chunk_rows: int = 65536):
    """fp32 SGEMM + top-k, the formulation of src/trainer.py:246-257 (logits = q @ base.T; topk).
    Used as the timed CPU baseline; row-chunked so the [B, N] logits never exceed a few hundred MB."""
    import torch

    qt = torch.from_numpy(np.ascontiguousarray(q, np.float32))
    best_v = None
    best_i = None
    for lo in range(0, db.shape[0], chunk_rows):
        blk = torch.from_numpy(np.ascontiguousarray(db[lo:lo + chunk_rows], np.float32))
        logits = qt @ blk.t()
        v, i = logits.topk(min(k, logits.shape[1]), dim=1)
        i = i + lo
        if best_v is None:
            best_v, best_i = v, i
        else:
            cv = torch.cat([best_v, v], 1)
            ci = torch.cat([best_i, i], 1)
            v2, sel = cv.topk(min(k, cv.shape[1]), dim=1)
            best_v, best_i = v2, torch.gather(ci, 1, sel)
    return best_v.numpy(), best_i.numpy()


def gather(base: np.ndarray, I: np.ndarray, perm: Optional[Sequence[int]] = None) -> np.ndarray:
    """base[I.reshape(-1)].reshape(B,k,-1), then feats[:, perm, :]  (src/trainer.py:215-219).
    Label -1 gathers zeros (the reference never produces it: k <= ntotal there)."""
    B, k = I.shape
    safe = np.where(I < 0, 0, I)
    out = np.asarray(base)[safe.reshape(-1)].reshape(B, k, -1).copy()
    out[I < 0] = 0
    if perm is not None:
        out = out[:, np.asarray(perm), :]
    return out


def weighted_pool(base: np.ndarray, I: np.ndarray, W: np.ndarray) -> np.ndarray:
    """out[b,h,:] = sum_j W[b,h,j] * base[I[b,j],:]  -- einsum('bhij,bhjd->bhid') with i = 1 of
    src/model/model.py:73, float64 accumulate."""
    feats = gather(np.asarray(base, np.float64), I)
    return np.einsum("bhj,bjd->bhd", np.asarray(W, np.float64), feats)


def softmax_weights(D: np.ndarray, tau: float) -> np.ndarray:
    """softmax over the k neighbours of tau * D  (shape of src/model/model.py:69-71), H = 1."""
    z = np.asarray(D, np.float64) * tau
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(axis=1, keepdims=True))[:, None, :]


def retrieved_features(feature: np.ndarray, image_base: np.ndarray, text_base: np.ndarray,
                       topk: int = 16, perm: Optional[Sequence[int]] = None, normalize: bool = True):
    """get_retrieved_features, use_faiss=True branch (src/trainer.py:198-230): normalise the
    query (:206), search both bases, gather, permute the image neighbours with one shared perm."""
    f = np.asarray(feature, np.float64)
    if normalize:
        f = f / np.linalg.norm(f, axis=1, keepdims=True)
    _, Ii = search(image_base, f, topk, "ip")
    _, It = search(text_base, f, topk, "ip")
    return gather(image_base, Ii, perm), gather(text_base, It, None), Ii, It


# ------------------------------------------------------------------------------------ metrics
def target_ranks(Q: np.ndarray, G: np.ndarray, target: np.ndarray,
                 exclude: Optional[np.ndarray] = None) -> np.ndarray:
    """rank[q] = #{g not in {target, exclude} : (s_g, -g) > (s_t, -t)}, float64 scores."""
    s = np.asarray(Q, np.float64) @ np.asarray(G, np.float64).T
    nq, ng = s.shape
    t = np.asarray(target, np.int64)
    st = s[np.arange(nq), t][:, None]
    gi = np.arange(ng)[None, :]
    better = (s > st) | ((s == st) & (gi < t[:, None]))
    better[np.arange(nq), t] = False
    if exclude is not None:
        better[np.arange(nq), np.asarray(exclude, np.int64)] = False
    return better.sum(1).astype(np.int64)


def metrics_cirr(image_features, ref_features, reference_names, index_names, target_names) -> Dict[str, float]:
    """get_metrics_cirr (src/eval_utils.py:1040-1067): rank of the target among the gallery with
    the query's own reference image removed (:1052-1056); R@k in percent (:1064-1065)."""
    names = [os.path.basename(n) for n in index_names]
    pos = {n: i for i, n in enumerate(names)}
    tgt = np.array([pos[n] for n in target_names], np.int64)
    ref = np.array([pos[n] for n in reference_names], np.int64)
    r = target_ranks(ref_features, image_features, tgt, ref)
    return {f"recall_R@{k}": float(np.sum(r < k)) / len(r) * 100 for k in [1, 5, 10, 50, 100]}


def cirr_testoutput(image_features, ref_features, reference_names, index_names, id_names) -> Dict:
    """get_cirr_testoutput (src/eval_utils.py:1070-1087): gallery names in ascending 1 - q.g order
    (float64 here, ties by lower gallery row), the query's own reference image removed (:1076-1080;
    names are matched raw, no basename), the first 50 kept, ".png" stripped (:1082-1086). Raises
    like the reference (IndexError) when fewer than 50 names remain."""
    s = np.asarray(ref_features, np.float64) @ np.asarray(image_features, np.float64).T
    pos = {n: i for i, n in enumerate(index_names)}
    out: Dict = {"version": "rc2", "metric": "recall"}
    for ind in range(len(id_names)):
        order = np.lexsort((np.arange(s.shape[1]), -s[ind]))
        row = [int(j) for j in order if j != pos[reference_names[ind]]]
        pid = id_names[ind]
        out[str(pid.item() if hasattr(pid, "item") else pid)] = [index_names[row[t]].replace(".png", "") for t in range(50)]
    return out


def metrics_fashion(image_features, ref_features, target_names, answer_names) -> Dict[str, float]:
    """get_metrics_fashion (src/eval_utils.py:1025-1037)."""
    pos = {n: i for i, n in enumerate(target_names)}
    tgt = np.array([pos[n] for n in answer_names], np.int64)
    r = target_ranks(ref_features, image_features, tgt)
    return {f"R@{k}": float(np.sum(r < k)) / len(r) * 100 for k in [1, 5, 10, 50, 100]}


def metrics_coco(image_features, ref_features, logit_scale: float = 1.0) -> Dict[str, float]:
    """get_metrics_coco (src/eval_utils.py:1008-1022): rank of the diagonal, both directions."""
    out: Dict[str, float] = {}
    n = len(ref_features)
    diag = np.arange(n, dtype=np.int64)
    for name, (A, Bm) in {"image_to_ref": (image_features, ref_features),
                          "ref_to_image": (ref_features, image_features)}.items():
        preds = target_ranks(A, Bm, diag)
        out[f"{name}_mean_rank"] = preds.mean() + 1
        out[f"{name}_median_rank"] = np.floor(np.median(preds)) + 1
        for k in [1, 5, 10, 50, 100]:
            out[f"{name}_R@{k}"] = float(np.mean(preds < k))
    return out


def metrics_imgnet(query_features, image_features, query_labels, target_labels) -> Dict[str, float]:
    """get_metrics_imgnet (src/eval_utils.py:1090-1134): hits among the top-k by label, recall =
    hits / (num_relevant + 1e-5) (:1115), precision = hits / k (:1116), means over queries."""
    ks = [1, 5, 10, 50, 100, 200]
    ql = np.asarray(query_labels, np.int64)
    tl = np.asarray(target_labels, np.int64)
    _, I = search(image_features, query_features, max(ks), "ip")
    hit = (tl[np.where(I < 0, 0, I)] == ql[:, None]) & (I >= 0)
    num_rel = np.bincount(tl, minlength=int(max(tl.max(), ql.max())) + 1)[ql].astype(np.float32)
    out: Dict[str, float] = {}
    for k in ks:
        c = hit[:, :k].sum(1).astype(np.float32)
        out[f"Real2Sketch_R@{k}"] = float(np.mean(c / (num_rel + np.float32(1e-5))))
        out[f"Real2Sketch_P@{k}"] = float(np.mean(c / np.float32(min(k, len(tl)))))
    return out


# ------------------------------------------------------------------------------------ comparison
def compare_topk(D_ref, I_ref, D_got, I_got, db=None, q=None, metric="ip",
                 tie_gap: float = 1e-5, d_tol: float = 1e-4) -> Dict[str, float]:
    """north_star parity rule: identical index sets except near-ties (score gap < tie_gap),
    distances within d_tol absolute. With db/q given, an index mismatch is excused only if the
    exact (float64) scores of the swapped rows differ by < tie_gap."""
    D_ref, I_ref, D_got, I_got = map(np.asarray, (D_ref, I_ref, D_got, I_got))
    assert D_ref.shape == D_got.shape and I_ref.shape == I_got.shape, (D_ref.shape, D_got.shape)
    B, k = I_ref.shape
    exact_rows = int(np.sum(np.all(I_ref == I_got, axis=1)))
    bad_sets = 0
    for b in np.nonzero(np.any(I_ref != I_got, axis=1))[0]:
        a, g = set(I_ref[b].tolist()), set(I_got[b].tolist())
        if a == g:
            # same set, different order: only legal between near-tied scores
            diff = np.nonzero(I_ref[b] != I_got[b])[0]
            if np.max(np.abs(D_ref[b, diff].astype(np.float64) - D_got[b, diff])) >= max(tie_gap, d_tol):
                bad_sets += 1
            continue
        only_ref, only_got = sorted(a - g), sorted(g - a)
        if db is None or len(only_ref) != len(only_got) or -1 in only_got or -1 in only_ref:
            bad_sets += 1
            continue
        qq = np.asarray(q[b], np.float64)
        def sc(ids):
            x = np.asarray(db[ids], np.float64)
            return x @ qq if metric == "ip" else -((x - qq) ** 2).sum(1)
        kth = np.sort(sc(I_ref[b][I_ref[b] >= 0]))[0]
        if np.any(np.abs(sc(only_got) - kth) >= tie_gap) or np.any(np.abs(sc(only_ref) - kth) >= tie_gap):
            bad_sets += 1
    # sorted per row, so a legal near-tie swap does not count; paddings (+-FLT_MAX) cancel exactly
    max_d = 0.0
    if B and k:
        max_d = float(np.max(np.abs(np.sort(D_ref, 1).astype(np.float64) - np.sort(D_got, 1).astype(np.float64))))
    return {"rows": B, "rows_identical": exact_rows, "bad_rows": bad_sets, "max_abs_D": max_d,
            "ok": bad_sets == 0 and max_d <= d_tol}
