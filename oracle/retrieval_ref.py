"""Oracle for Path 2 (retrieval scoring).  TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Restates, in numpy, what the reference computes in
  * epoch_original.py:115-161  ``itm_eval``  (upstream, loop + np.where)
  * epoch.py:219-244           ``itm_eval``  (fork, np.isin formulation)
  * epoch_original.py:77-111   ``epoch_test`` from the embeddings onward
    (normalise, exp(logit_scale) * I @ T^T, per-row top-128 kept / -100 fill)

Rank definition.  The reference ranks with ``np.argsort(score)[::-1]`` which is an
unstable sort: the order among EQUAL scores is implementation-defined.  The oracle
(and the CUDA path) use the deterministic equivalent named in BASELINE.json
("ties broken by index"):

    rank(i -> c) = #{j : s_j > s_c} + #{j < c : s_j == s_c}

i.e. the position of c in a stable descending sort (lower index first among ties).
On tie-free rows this equals the reference's rank exactly; with ties it is one of
the orders numpy may produce, and recall@{1,5,10} after the top-128/-100 fill is
invariant (all tied -100 entries sit at rank >= 128).
"""
from __future__ import annotations

import numpy as np

RESULT_KEYS = ("txt_r1", "txt_r5", "txt_r10", "txt_r_mean",
               "img_r1", "img_r5", "img_r10", "img_r_mean", "r_mean")


def stable_desc_rank(row: np.ndarray, c: int) -> int:
    """Position of column ``c`` in a stable descending sort of ``row``."""
    s = row[c]
    return int(np.count_nonzero(row > s) + np.count_nonzero(row[:c] == s))


def ranks_i2t(scores_i2t: np.ndarray, img2txt) -> np.ndarray:
    """epoch_original.py:117-128 / epoch.py:222-226: best (minimum) rank over the image's GT captions."""
    n = scores_i2t.shape[0]
    out = np.zeros(n, dtype=np.int32)
    for i in range(n):
        row = scores_i2t[i]
        out[i] = min(stable_desc_rank(row, int(c)) for c in img2txt[i])
    return out


def ranks_t2i(scores_t2i: np.ndarray, txt2img) -> np.ndarray:
    """epoch_original.py:136-141 / epoch.py:231-235: rank of the caption's GT image."""
    n = scores_t2i.shape[0]
    out = np.zeros(n, dtype=np.int32)
    for t in range(n):
        out[t] = stable_desc_rank(scores_t2i[t], int(txt2img[t]))
    return out


def ranks_vectorised(scores: np.ndarray, gt_ptr: np.ndarray, gt_idx: np.ndarray) -> np.ndarray:
    """Same definition as above for CSR ground truth, vectorised per row (fast enough for 5000x25000)."""
    n = scores.shape[0]
    out = np.empty(n, dtype=np.int32)
    for r in range(n):
        row = scores[r]
        best = None
        for c in gt_idx[gt_ptr[r]:gt_ptr[r + 1]]:
            c = int(c)
            s = row[c]
            k = int(np.count_nonzero(row > s) + np.count_nonzero(row[:c] == s))
            best = k if best is None or k < best else best
        out[r] = best
    return out


def rank_bounds(scores: np.ndarray, gt_lists) -> tuple:
    """(optimistic, pessimistic) ranks per row: every tie resolved for / against the ground truth.

    Any argsort -- stable or not, e.g. numpy's SIMD introsort whose tie order is machine-dependent -- yields a rank
    inside these bounds, so the reference's recall on tied data must lie between the two recalls.
    """
    n = scores.shape[0]
    lo, hi = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
    for r in range(n):
        row = scores[r]
        best_lo, best_hi = None, None
        for c in gt_lists[r]:
            s = row[int(c)]
            g = int(np.count_nonzero(row > s))
            e = int(np.count_nonzero(row == s))
            best_lo = g if best_lo is None else min(best_lo, g)
            best_hi = g + e - 1 if best_hi is None else min(best_hi, g + e - 1)
        lo[r], hi[r] = best_lo, best_hi
    return lo, hi


def recall_dict(ranks_img: np.ndarray, ranks_txt: np.ndarray) -> dict:
    """epoch_original.py:131-161 / epoch.py:227-244: recall@1/5/10 in percent and the three means.

    ``txt_*`` keys are image->text retrieval, ``img_*`` keys are text->image (reference naming).
    """
    def r_at(ranks, k):
        return 100.0 * int(np.count_nonzero(ranks < k)) / len(ranks)
    tr1, tr5, tr10 = (r_at(ranks_img, k) for k in (1, 5, 10))
    ir1, ir5, ir10 = (r_at(ranks_txt, k) for k in (1, 5, 10))
    tr_mean = (tr1 + tr5 + tr10) / 3
    ir_mean = (ir1 + ir5 + ir10) / 3
    return {"txt_r1": tr1, "txt_r5": tr5, "txt_r10": tr10, "txt_r_mean": tr_mean,
            "img_r1": ir1, "img_r5": ir5, "img_r10": ir10, "img_r_mean": ir_mean,
            "r_mean": (tr_mean + ir_mean) / 2}


def itm_eval_ref(scores_i2t, scores_t2i, txt2img, img2txt) -> dict:
    """Signature of epoch.py:219 / epoch_original.py:115."""
    return recall_dict(ranks_i2t(np.asarray(scores_i2t), img2txt),
                       ranks_t2i(np.asarray(scores_t2i), txt2img))


def l2_normalise(x: np.ndarray) -> np.ndarray:
    """epoch_original.py:78,84,92: x / x.norm(dim=1, keepdim=True) (no eps)."""
    return x / np.sqrt((x.astype(np.float64) ** 2).sum(axis=1, keepdims=True)).astype(x.dtype)


def sims_ref(img_embeds: np.ndarray, txt_embeds: np.ndarray, logit_scale_exp: float) -> np.ndarray:
    """epoch_original.py:94: exp(logit_scale) * image_embeds @ text_embeds.t() (scale applied to the image rows first)."""
    return (np.float32(logit_scale_exp) * img_embeds) @ txt_embeds.T


def topk_fill_ref(sims: np.ndarray, k: int = 128, fill: float = -100.0) -> np.ndarray:
    """epoch_original.py:95-99 (and 101-105 on the transpose): keep each row's top-k, everything else := fill.

    torch.topk picks, among equal values, an implementation-defined subset; the oracle keeps the
    lower indices first (stable), which only matters when the k-th value is tied.
    """
    out = np.full_like(sims, fill)
    kk = min(k, sims.shape[1])
    for r in range(sims.shape[0]):
        idx = np.argsort(-sims[r], kind="stable")[:kk]
        out[r, idx] = sims[r, idx]
    return out


def epoch_test_ref(img_embeds, txt_embeds, logit_scale_exp=float(np.exp(np.log(1 / 0.07))), k=128):
    """epoch_original.py:92-111 from (already row-normalised) embeddings: returns (score_i2t[I,T], score_t2i[T,I])."""
    sims = sims_ref(img_embeds, txt_embeds, logit_scale_exp)
    return topk_fill_ref(sims, k), topk_fill_ref(np.ascontiguousarray(sims.T), k)


def flickr_maps(n_img: int, caps_per_img: int = 5):
    """flickr30k_dataset.py:110-118: captions of image i are the contiguous block [C*i, C*i+C)."""
    img2txt = {i: list(range(caps_per_img * i, caps_per_img * (i + 1))) for i in range(n_img)}
    txt2img = {t: t // caps_per_img for t in range(n_img * caps_per_img)}
    return txt2img, img2txt


def synthetic_retrieval(n_img, caps_per_img, dim, seed=0, corr=0.15, dtype=np.float32):
    """correlated gaussian embeddings, rows L2-normalised."""
    rng = np.random.default_rng(seed)
    img = rng.standard_normal((n_img, dim)).astype(dtype)
    txt = rng.standard_normal((n_img * caps_per_img, dim)).astype(dtype)
    txt += dtype(corr) * np.repeat(img, caps_per_img, axis=0)
    return l2_normalise(img), l2_normalise(txt)


# ---- nearest-neighbour caption lookup (distill.py:89-95) -------------------------------------------------------------
def nearest_problem(tag: str):
    """Seeded (query, bank) pair: 'small' has an exact duplicate row (first index must win) and an all-zero row;
    'mid' is 100 queries against 3000 rows of BERT-like 768-d embeddings."""
    Q, T, D = {"small": (7, 50, 12), "mid": (100, 3000, 768)}[tag]
    rng = np.random.default_rng(5 if tag == "small" else 6)
    bank = (rng.standard_normal((T, D)) * 0.5253 - 0.0094).astype(np.float32)
    query = (bank[rng.integers(0, T, Q)] + 0.3 * rng.standard_normal((Q, D))).astype(np.float32)
    if tag == "small":
        bank[10] = bank[3]
        query[0] = 2.0 * bank[3]
        bank[20] = 0.0
    return query, bank


def nearest_neighbor_ref(query: np.ndarray, bank: np.ndarray) -> np.ndarray:
    """distill.py:89-95 restated: per query argmax of sklearn's cosine_similarity (rows L2-normalised, all-zero rows kept
    as zeros: sklearn.preprocessing.normalize divides by 1 when the norm is 0), first index on ties (np.argmax)."""
    def unit(x):
        x = np.asarray(x, dtype=np.float64)
        n = np.sqrt((x * x).sum(axis=1, keepdims=True))
        n[n == 0.0] = 1.0
        return x / n
    return np.argmax(unit(query) @ unit(bank).T, axis=1).astype(np.int32)


def cosine_matrix(query: np.ndarray, bank: np.ndarray) -> np.ndarray:
    """sklearn.metrics.pairwise.cosine_similarity restated in float64 (zero rows stay zero)."""
    def unit(x):
        x = np.asarray(x, dtype=np.float64)
        n = np.sqrt((x * x).sum(axis=1, keepdims=True))
        n[n == 0.0] = 1.0
        return x / n
    return unit(query) @ unit(bank).T


# ---- plain-C restatement of the integer part (oracle/c/itm_eval_ref.c), an independent cross-check of the numpy one -----
def c_oracle():
    """ctypes handle of oracle/_build/libitm_port.so (built by `make -C oracle/c`, which __graft_entry__.build() runs)."""
    import ctypes
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "_build", "libitm_port.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-s", "-C", os.path.join(here, "c")], check=True)
    lib = ctypes.CDLL(so)
    P = ctypes.c_void_p
    lib.itm_ranks_rows.argtypes = [P, ctypes.c_int, ctypes.c_int, P, P, P]
    lib.itm_ranks_rows.restype = None
    lib.itm_result.argtypes = [P, ctypes.c_int, P, ctypes.c_int, P]
    lib.itm_result.restype = None
    return lib


def itm_eval_c(scores_i2t, scores_t2i, txt2img, img2txt, return_ranks=False):
    """itm_eval through the C restatement: same inputs as the reference (numpy matrices + the dataset's dict maps)."""
    lib = c_oracle()
    s1 = np.ascontiguousarray(scores_i2t, dtype=np.float32)
    s2 = np.ascontiguousarray(scores_t2i, dtype=np.float32)
    I, T = s1.shape
    ptr = np.zeros(I + 1, dtype=np.int32)
    lists = [np.atleast_1d(np.asarray(img2txt[i], dtype=np.int32)) for i in range(I)]
    ptr[1:] = np.cumsum([len(l) for l in lists])
    idx = np.concatenate(lists).astype(np.int32)
    t2i = np.asarray([txt2img[t] for t in range(T)], dtype=np.int32)
    r1, r2 = np.empty(I, dtype=np.int32), np.empty(T, dtype=np.int32)
    a = lambda x: x.ctypes.data
    lib.itm_ranks_rows(a(s1), I, T, a(ptr), a(idx), a(r1))
    one = np.arange(T + 1, dtype=np.int32)
    lib.itm_ranks_rows(a(s2), T, I, a(one), a(t2i), a(r2))
    out = np.zeros(9, dtype=np.float64)
    lib.itm_result(a(r1), I, a(r2), T, a(out))
    res = {k: float(v) for k, v in zip(RESULT_KEYS, out)}
    return (res, r1, r2) if return_ranks else res
