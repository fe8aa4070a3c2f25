"""torch-tensor front end of the C ABI: argument checking, output allocation, stream plumbing.

torch is used for device memory and streams only; all arithmetic happens in libvldd_b200.so.  Every function
raises on non-CUDA input -- there is no CPU path (BASELINE.json north_star: "no CPU fallback").
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import check, lib

LOGIT_SCALE_EVAL = float(np.exp(np.log(1 / 0.07)))      # epoch_original.py:70,94 / networks.py:878  (= 14.2857...)
LOGIT_SCALE_UPSTREAM = float(np.log(1 / 0.07))          # distill_original.py:103,430 (used un-exponentiated)


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _req(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (this library has no CPU path), got device {t.device}")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _flat(t: torch.Tensor) -> torch.Tensor:
    """Flat-parameter convention: [P] or the DataParallel [1, P] slice (reparam_module.py:149 squeezes it)."""
    return t.reshape(-1) if isinstance(t, torch.Tensor) else t


def _scalar(v, device, name: str) -> torch.Tensor:
    """Scalars the reference keeps as tensors (syn_lr, logit scale) stay on the device: no host sync."""
    if isinstance(v, torch.Tensor):
        t = v.detach()
        if t.numel() != 1:
            raise ValueError(f"{name}: expected one element, got shape {tuple(t.shape)}")
        return t.to(device=device, dtype=torch.float32).reshape(1).contiguous()
    return torch.full((1,), float(v), dtype=torch.float32, device=device)


def head_numel(dt: int, d: int) -> int:
    return d * dt + d + d * d + d + d + d


# --------------------------------------------------------------------------------------------------
# streaming
# --------------------------------------------------------------------------------------------------
def flat_sgd_step(theta: torch.Tensor, grad: torch.Tensor, lr, out: torch.Tensor | None = None) -> torch.Tensor:
    """theta - lr * grad  (distill.py:582-583)."""
    theta = _req(_flat(theta), "theta")
    grad = _req(_flat(grad), "grad")
    if theta.shape != grad.shape:
        raise ValueError(f"theta {tuple(theta.shape)} and grad {tuple(grad.shape)} differ")
    out = torch.empty_like(theta) if out is None else _req(out, "out")
    lr_t = _scalar(lr, theta.device, "lr")
    check(lib().vldd_flat_sgd_step(_ptr(theta), _ptr(grad), _ptr(lr_t), _ptr(out), theta.numel(), _stream()),
          "flat_sgd_step")
    return out


def match_loss(theta_K: torch.Tensor, theta_tgt: torch.Tensor, theta_0: torch.Tensor) -> torch.Tensor:
    """Returns a device tensor [num, den, num/den]  (distill.py:588-598)."""
    a, b, c = (_req(_flat(t), n) for t, n in ((theta_K, "theta_K"), (theta_tgt, "theta_tgt"), (theta_0, "theta_0")))
    if not (a.shape == b.shape == c.shape):
        raise ValueError("match_loss: shape mismatch")
    out = torch.empty(3, dtype=torch.float32, device=a.device)
    scratch = torch.zeros(lib().vldd_match_loss_scratch_bytes(), dtype=torch.uint8, device=a.device)
    check(lib().vldd_match_loss_fwd(_ptr(a), _ptr(b), _ptr(c), a.numel(), _ptr(out), _ptr(scratch), _stream()),
          "match_loss_fwd")
    return out


def match_loss_bwd(theta_K: torch.Tensor, theta_tgt: torch.Tensor, num_den: torch.Tensor, gout=None) -> torch.Tensor:
    a, b = _req(_flat(theta_K), "theta_K"), _req(_flat(theta_tgt), "theta_tgt")
    nd = _req(num_den, "num_den")
    g = None if gout is None else _scalar(gout, a.device, "gout")
    out = torch.empty_like(a)
    check(lib().vldd_match_loss_bwd(_ptr(a), _ptr(b), _ptr(nd), _ptr(g), _ptr(out), a.numel(), _stream()), "match_loss_bwd")
    return out


def momentum_sgd_(param: torch.Tensor, grad: torch.Tensor, buf: torch.Tensor, lr: float, momentum: float, first: bool) -> None:
    """In-place torch.optim.SGD(momentum, dampening=0) step (distill.py:233-241, 611-613)."""
    for t, n in ((param, "param"), (grad, "grad"), (buf, "buf")):
        _req(t, n)
        if not t.is_contiguous():
            raise ValueError(f"{n} must be contiguous for the in-place update")
    check(lib().vldd_momentum_sgd(_ptr(param), _ptr(grad), _ptr(buf), float(lr), float(momentum), int(bool(first)),
                                  param.numel(), _stream()), "momentum_sgd")


# --------------------------------------------------------------------------------------------------
# retrieval
# --------------------------------------------------------------------------------------------------
def maps_to_arrays(txt2img, img2txt, n_img: int, n_txt: int):
    """dict[int->int], dict[int->list[int]] (flickr30k_dataset.py:110-118) -> (txt2img[T], ptr[I+1], idx[]) int32."""
    t2i = np.empty(n_txt, dtype=np.int32)
    if isinstance(txt2img, dict):
        for t in range(n_txt):
            t2i[t] = txt2img[t]
    else:
        t2i[:] = np.asarray(txt2img, dtype=np.int32)[:n_txt]
    ptr = np.zeros(n_img + 1, dtype=np.int32)
    lists = [np.atleast_1d(np.asarray(img2txt[i], dtype=np.int32)) for i in range(n_img)]
    for i, l in enumerate(lists):
        ptr[i + 1] = ptr[i] + len(l)
    idx = np.concatenate(lists).astype(np.int32) if lists else np.zeros(0, dtype=np.int32)
    # the kernels index the score matrices with these: refuse maps that point outside them (numpy would raise IndexError)
    if n_txt and (t2i.min() < 0 or t2i.max() >= n_img):
        raise IndexError(f"txt2img refers to image {int(t2i.max() if t2i.max() >= n_img else t2i.min())}, but there are {n_img} images")
    if idx.size and (idx.min() < 0 or idx.max() >= n_txt):
        raise IndexError(f"img2txt refers to caption {int(idx.max() if idx.max() >= n_txt else idx.min())}, but there are {n_txt} captions")
    return t2i, ptr, idx


def ranks_from_scores(scores_i2t, scores_t2i, txt2img: torch.Tensor, img2txt_ptr: torch.Tensor, img2txt_idx: torch.Tensor):
    """Device score matrices -> (ranks_i2t[I], ranks_t2i[T]) int32 device tensors."""
    s1 = None if scores_i2t is None else _req(scores_i2t, "scores_i2t")
    s2 = None if scores_t2i is None else _req(scores_t2i, "scores_t2i")
    ref = s1 if s1 is not None else s2
    if ref is None:
        raise ValueError("need at least one score matrix")
    n_img, n_txt = (s1.shape if s1 is not None else s2.shape[::-1])
    if s1 is not None and s2 is not None and tuple(s2.shape) != (n_txt, n_img):
        raise ValueError(f"scores_t2i must be [{n_txt},{n_img}], got {tuple(s2.shape)}")
    t2i = _req(txt2img, "txt2img", torch.int32)
    ptr = _req(img2txt_ptr, "img2txt_ptr", torch.int32)
    idx = _req(img2txt_idx, "img2txt_idx", torch.int32)
    r1 = torch.empty(n_img, dtype=torch.int32, device=ref.device)
    r2 = torch.empty(n_txt, dtype=torch.int32, device=ref.device)
    check(lib().vldd_ranks_from_scores(_ptr(s1), _ptr(s2), n_img, n_txt, _ptr(t2i), _ptr(ptr), _ptr(idx), _ptr(r1),
                                       _ptr(r2), _stream()), "ranks_from_scores")
    return (r1 if s1 is not None else None), (r2 if s2 is not None else None)


def ranks_cols(scores_i2t: torch.Tensor, txt2img: torch.Tensor) -> torch.Tensor:
    """Text->image ranks read column-wise from the image->text matrix [I, T] (T may be a caption shard)."""
    s = _req(scores_i2t, "scores_i2t")
    I, T = s.shape
    out = torch.empty(T, dtype=torch.int32, device=s.device)
    check(lib().vldd_ranks_cols(_ptr(s), I, T, _ptr(_req(txt2img, "txt2img", torch.int32)), _ptr(out), _stream()), "ranks_cols")
    return out


def rank_best_gt(scores: torch.Tensor, col_offset: int, gt_ptr: torch.Tensor, gt_idx: torch.Tensor):
    """Best local ground-truth candidate per row of a column shard: (score[rows] f32, global index[rows] i32)."""
    s = _req(scores, "scores")
    rows, cols = s.shape
    bs = torch.empty(rows, dtype=torch.float32, device=s.device)
    bi = torch.empty(rows, dtype=torch.int32, device=s.device)
    check(lib().vldd_rank_best_gt(_ptr(s), rows, cols, int(col_offset), _ptr(_req(gt_ptr, "gt_ptr", torch.int32)),
                                  _ptr(_req(gt_idx, "gt_idx", torch.int32)), _ptr(bs), _ptr(bi), _stream()), "rank_best_gt")
    return bs, bi


def rank_count(scores: torch.Tensor, col_offset: int, thr_score: torch.Tensor, thr_idx: torch.Tensor) -> torch.Tensor:
    """Per row: number of local columns ranked ahead of the (global) threshold candidate."""
    s = _req(scores, "scores")
    rows, cols = s.shape
    out = torch.empty(rows, dtype=torch.int32, device=s.device)
    check(lib().vldd_rank_count(_ptr(s), rows, cols, int(col_offset), _ptr(_req(thr_score, "thr_score")),
                                _ptr(_req(thr_idx, "thr_idx", torch.int32)), _ptr(out), _stream()), "rank_count")
    return out


def recall_counts(ranks: torch.Tensor) -> torch.Tensor:
    r = _req(ranks, "ranks", torch.int32)
    out = torch.empty(3, dtype=torch.int32, device=r.device)
    check(lib().vldd_recall_counts(_ptr(r), r.numel(), _ptr(out), _stream()), "recall_counts")
    return out


def sim_scores(img: torch.Tensor, txt: torch.Tensor, scale: float = LOGIT_SCALE_EVAL, want_t2i: bool = True):
    img, txt = _req(img, "img"), _req(txt, "txt")
    if img.shape[1] != txt.shape[1]:
        raise ValueError("embedding dims differ")
    I, T, D = img.shape[0], txt.shape[0], img.shape[1]
    s1 = torch.empty(I, T, dtype=torch.float32, device=img.device)
    s2 = torch.empty(T, I, dtype=torch.float32, device=img.device) if want_t2i else None
    check(lib().vldd_sim_scores(_ptr(img), _ptr(txt), I, T, D, float(scale), _ptr(s1), _ptr(s2), _stream()), "sim_scores")
    return s1, s2


def topk_fill(scores: torch.Tensor, k: int = 128, fill: float = -100.0) -> torch.Tensor:
    s = _req(scores, "scores")
    out = torch.empty_like(s)
    check(lib().vldd_topk_fill(_ptr(s), _ptr(out), s.shape[0], s.shape[1], int(k), float(fill), _stream()), "topk_fill")
    return out


def sim_rank(img: torch.Tensor, txt: torch.Tensor, txt2img: torch.Tensor, img2txt_ptr: torch.Tensor,
             img2txt_idx: torch.Tensor, scale: float = LOGIT_SCALE_EVAL, workspace: torch.Tensor | None = None):
    img, txt = _req(img, "img"), _req(txt, "txt")
    I, T, D = img.shape[0], txt.shape[0], img.shape[1]
    need = lib().vldd_sim_rank_workspace_bytes(I, T, D)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=img.device)
    r1 = torch.empty(I, dtype=torch.int32, device=img.device)
    r2 = torch.empty(T, dtype=torch.int32, device=img.device)
    check(lib().vldd_sim_rank(_ptr(img), _ptr(txt), I, T, D, float(scale), _ptr(_req(txt2img, "txt2img", torch.int32)),
                              _ptr(_req(img2txt_ptr, "img2txt_ptr", torch.int32)),
                              _ptr(_req(img2txt_idx, "img2txt_idx", torch.int32)), _ptr(r1), _ptr(r2), _ptr(workspace),
                              workspace.numel(), _stream()), "sim_rank")
    return r1, r2


def sim_rank_fused(img: torch.Tensor, txt: torch.Tensor, txt2img: torch.Tensor, img2txt_ptr: torch.Tensor,
                   img2txt_idx: torch.Tensor, scale: float = LOGIT_SCALE_EVAL, workspace: torch.Tensor | None = None):
    """Embeddings -> ranks of both directions with the ranking fused into the GEMM epilogue (no score matrix in HBM)."""
    img, txt = _req(img, "img"), _req(txt, "txt")
    I, T, D = img.shape[0], txt.shape[0], img.shape[1]
    idx = _req(img2txt_idx, "img2txt_idx", torch.int32)
    nnz = idx.numel()
    need = lib().vldd_sim_rank_fused_workspace_bytes(I, T, D, nnz)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=img.device)
    r1 = torch.empty(I, dtype=torch.int32, device=img.device)
    r2 = torch.empty(T, dtype=torch.int32, device=img.device)
    check(lib().vldd_sim_rank_fused(_ptr(img), _ptr(txt), I, T, D, float(scale), _ptr(_req(txt2img, "txt2img", torch.int32)),
                                    _ptr(_req(img2txt_ptr, "img2txt_ptr", torch.int32)), _ptr(idx), nnz, _ptr(r1), _ptr(r2),
                                    _ptr(workspace), workspace.numel(), _stream()), "sim_rank_fused")
    return r1, r2


class FusedRankShard:
    """Caption-shard form of sim_rank_fused (multi-GPU retrieval): phase A ``candidates()``, exchange, phase B ``count()``.

    img [I, D] replicated; txt_shard [T_r, D] = captions [lo, lo + T_r) of the full set; txt2img_shard int32 [T_r];
    img2txt CSR with GLOBAL caption ids.  The workspace is owned by the object and shared by both phases.
    """

    def __init__(self, img, txt_shard, lo: int, txt2img_shard, img2txt_ptr, img2txt_idx, scale: float = LOGIT_SCALE_EVAL):
        self.img, self.txt = _req(img, "img"), _req(txt_shard, "txt_shard")
        self.t2i = _req(txt2img_shard, "txt2img_shard", torch.int32)
        self.ptr, self.idx = _req(img2txt_ptr, "img2txt_ptr", torch.int32), _req(img2txt_idx, "img2txt_idx", torch.int32)
        self.lo, self.scale = int(lo), float(scale)
        self.I, self.T, self.D = self.img.shape[0], self.txt.shape[0], self.img.shape[1]
        self.nnz = self.idx.numel()
        need = lib().vldd_sim_rank_fused_workspace_bytes(self.I, self.T, self.D, self.nnz)
        self.ws = torch.empty(need, dtype=torch.uint8, device=self.img.device)

    def candidates(self):
        """(score[I] f32, +inf where the image has no ground truth in this shard; global caption id[I] i32, -1 = none)."""
        cs = torch.empty(self.I, dtype=torch.float32, device=self.img.device)
        ci = torch.empty(self.I, dtype=torch.int32, device=self.img.device)
        check(lib().vldd_sim_rank_fused_candidates(_ptr(self.img), _ptr(self.txt), self.I, self.T, self.D, self.scale, _ptr(self.t2i),
                                                   _ptr(self.ptr), _ptr(self.idx), self.nnz, self.lo, _ptr(cs), _ptr(ci), _ptr(self.ws),
                                                   self.ws.numel(), _stream()), "sim_rank_fused_candidates")
        return cs, ci

    def count(self, thr_score: torch.Tensor, thr_idx_global: torch.Tensor, invalid_row_rank: int = 0):
        """(row_counts[I], ranks_t2i[T_r]) for the merged thresholds (score +inf / index -1: no ground truth anywhere)."""
        ts = _req(thr_score, "thr_score")
        ti = _req(thr_idx_global, "thr_idx_global", torch.int32)
        ts = torch.where(ti >= 0, ts, torch.full_like(ts, float("inf"))).contiguous()
        tl = (ti - self.lo).to(torch.int32).contiguous()
        rc = torch.empty(self.I, dtype=torch.int32, device=self.img.device)
        rt = torch.empty(self.T, dtype=torch.int32, device=self.img.device)
        check(lib().vldd_sim_rank_fused_count(_ptr(self.img), _ptr(self.txt), self.I, self.T, self.D, self.scale, _ptr(ts), _ptr(tl),
                                              self.nnz, int(invalid_row_rank), _ptr(rc), _ptr(rt), _ptr(self.ws), self.ws.numel(),
                                              _stream()), "sim_rank_fused_count")
        return rc, rt


RESULT_KEYS = ("txt_r1", "txt_r5", "txt_r10", "txt_r_mean", "img_r1", "img_r5", "img_r10", "img_r_mean", "r_mean")


def itm_eval_host(scores_i2t: np.ndarray, scores_t2i: np.ndarray, t2i: np.ndarray, ptr: np.ndarray, idx: np.ndarray,
                  want_ranks: bool = False):
    """Host numpy matrices in, result dict out -- the C-ABI drop-in under epoch.itm_eval."""
    if not torch.cuda.is_available():
        raise RuntimeError("itm_eval needs a CUDA device (this library has no CPU path)")
    s1 = np.ascontiguousarray(scores_i2t, dtype=np.float32)
    s2 = np.ascontiguousarray(scores_t2i, dtype=np.float32)
    I, T = s1.shape
    if s2.shape != (T, I):
        raise ValueError(f"scores_t2i must be [{T},{I}], got {s2.shape}")
    res = np.zeros(9, dtype=np.float64)
    r1 = np.empty(I, dtype=np.int32) if want_ranks else None
    r2 = np.empty(T, dtype=np.int32) if want_ranks else None
    vp = lambda a: C.c_void_p(0 if a is None else a.ctypes.data)
    check(lib().vldd_itm_eval_host(vp(s1), vp(s2), I, T, vp(t2i), vp(ptr), vp(idx), vp(r1), vp(r2), vp(res), _stream()),
          "itm_eval_host")
    out = {k: float(v) for k, v in zip(RESULT_KEYS, res)}
    return (out, r1, r2) if want_ranks else out


# --------------------------------------------------------------------------------------------------
# head / InfoNCE / unroll
# --------------------------------------------------------------------------------------------------
def proj_head_forward(theta: torch.Tensor, Y: torch.Tensor, d: int, mask: torch.Tensor | None = None,
                      normalise: bool = False) -> torch.Tensor:
    theta = _req(_flat(theta), "theta")
    Y = _req(Y, "Y")
    rows, dt = Y.shape
    if theta.numel() != head_numel(dt, d):
        raise ValueError(f"theta has {theta.numel()} elements, expected {head_numel(dt, d)} for dt={dt}, d={d}")
    mask = None if mask is None else _req(mask, "mask")
    out = torch.empty(rows, d, dtype=torch.float32, device=Y.device)
    ws = torch.empty(lib().vldd_proj_head_workspace_bytes(rows, dt, d), dtype=torch.uint8, device=Y.device)
    z, zn = (None, out) if normalise else (out, None)
    check(lib().vldd_proj_head_forward(_ptr(theta), _ptr(Y), _ptr(mask), rows, dt, d, _ptr(z), _ptr(zn), _ptr(ws),
                                       ws.numel(), _stream()), "proj_head_forward")
    return out


def contrastive_step(theta: torch.Tensor, Y: torch.Tensor, U: torch.Tensor, scale, mask: torch.Tensor | None = None):
    """loss and first-order grads of one step (config 2): returns dict(loss, g_theta, dY, dU, dscale)."""
    theta, Y, U = _req(_flat(theta), "theta"), _req(Y, "Y"), _req(U, "U")
    B, dt = Y.shape
    d = U.shape[1]
    if theta.numel() != head_numel(dt, d):
        raise ValueError(f"theta has {theta.numel()} elements, expected {head_numel(dt, d)}")
    dev = Y.device
    sc = _scalar(scale, dev, "scale")
    mask = None if mask is None else _req(mask, "mask")
    loss = torch.empty(1, device=dev)
    g = torch.empty_like(theta)
    dY, dU, dsc = torch.empty_like(Y), torch.empty_like(U), torch.empty(1, device=dev)
    ws = torch.empty(lib().vldd_contrastive_step_workspace_bytes(B, dt, d), dtype=torch.uint8, device=dev)
    check(lib().vldd_contrastive_step(_ptr(theta), _ptr(Y), _ptr(U), _ptr(sc), _ptr(mask), B, dt, d, _ptr(loss), _ptr(g),
                                      _ptr(dY), _ptr(dU), _ptr(dsc), _ptr(ws), ws.numel(), _stream()), "contrastive_step")
    return dict(loss=loss[0], g_theta=g, dY=dY, dU=dU, dscale=dsc[0])


def clip_loss(theta: torch.Tensor, Y: torch.Tensor, U: torch.Tensor, scale=LOGIT_SCALE_EVAL, mask: torch.Tensor | None = None):
    """networks.py:866-889 from the encoder outputs on: dict(loss, top1[2] int32, g_theta, dY, dU, dscale).

    top1 = (#image rows whose best caption is their own, #captions whose best image is their own); the reference's
    ``acc`` is ``top1.sum() / 2``.
    """
    theta, Y, U = _req(_flat(theta), "theta"), _req(Y, "Y"), _req(U, "U")
    B, dt = Y.shape
    d = U.shape[1]
    if U.shape[0] != B:
        raise ValueError("image and text batches differ in size")
    if theta.numel() != head_numel(dt, d):
        raise ValueError(f"theta has {theta.numel()} elements, expected {head_numel(dt, d)}")
    dev = Y.device
    sc = _scalar(scale, dev, "scale")
    mask = None if mask is None else _req(mask, "mask")
    loss = torch.empty(1, device=dev)
    top1 = torch.empty(2, dtype=torch.int32, device=dev)
    g = torch.empty_like(theta)
    dY, dU, dsc = torch.empty_like(Y), torch.empty_like(U), torch.empty(1, device=dev)
    ws = torch.empty(lib().vldd_contrastive_step_workspace_bytes(B, dt, d), dtype=torch.uint8, device=dev)
    check(lib().vldd_clip_loss(_ptr(theta), _ptr(Y), _ptr(U), _ptr(sc), _ptr(mask), B, dt, d, _ptr(loss), _ptr(top1), _ptr(g),
                               _ptr(dY), _ptr(dU), _ptr(dsc), _ptr(ws), ws.numel(), _stream()), "clip_loss")
    return dict(loss=loss[0], top1=top1, g_theta=g, dY=dY, dU=dU, dscale=dsc[0])


def nearest_rows(query: torch.Tensor, bank: torch.Tensor, return_cos: bool = False):
    """Index of the most cosine-similar bank row per query row, first index on ties (distill.py:89-95)."""
    q, b = _req(query, "query"), _req(bank, "bank")
    if q.dim() != 2 or b.dim() != 2 or q.shape[1] != b.shape[1]:
        raise ValueError(f"query {tuple(q.shape)} and bank {tuple(b.shape)} must be 2-D with equal width")
    Q, D = q.shape
    T = b.shape[0]
    idx = torch.empty(Q, dtype=torch.int32, device=q.device)
    cos = torch.empty(Q, dtype=torch.float32, device=q.device) if return_cos else None
    ws = torch.empty(max(lib().vldd_nearest_rows_workspace_bytes(Q, T, D), 1), dtype=torch.uint8, device=q.device)
    check(lib().vldd_nearest_rows(_ptr(q), _ptr(b), Q, T, D, _ptr(idx), _ptr(cos), _ptr(ws), ws.numel(), _stream()),
          "nearest_rows")
    return (idx, cos) if return_cos else idx


def _nce_args(xn, yn, scale):
    xn, yn = _req(xn, "xn"), _req(yn, "yn")
    if xn.dim() != 2 or xn.shape != yn.shape:
        raise ValueError(f"xn {tuple(xn.shape)} and yn {tuple(yn.shape)} must be equal 2-D shapes")
    B, d = xn.shape
    sc = _scalar(scale, xn.device, "scale")
    ws = torch.empty(max(lib().vldd_infonce_workspace_bytes(B, d), 1), dtype=torch.uint8, device=xn.device)
    return xn, yn, sc, B, d, ws


def infonce_grad(xn: torch.Tensor, yn: torch.Tensor, scale):
    """Bidirectional InfoNCE on row-normalised features (distill.py:548-551): dict(loss, dxn, dyn, dscale)."""
    xn, yn, sc, B, d, ws = _nce_args(xn, yn, scale)
    loss = torch.empty(1, device=xn.device)
    dxn, dyn, dsc = torch.empty_like(xn), torch.empty_like(yn), torch.empty(1, device=xn.device)
    check(lib().vldd_infonce_grad(_ptr(xn), _ptr(yn), _ptr(sc), B, d, _ptr(loss), _ptr(dxn), _ptr(dyn), _ptr(dsc), _ptr(ws),
                                  ws.numel(), _stream()), "infonce_grad")
    return dict(loss=loss[0], dxn=dxn, dyn=dyn, dscale=dsc[0])


def infonce_hvp(xn: torch.Tensor, yn: torch.Tensor, scale, cx: torch.Tensor, cy: torch.Tensor, cs):
    """Hessian of the InfoNCE loss applied to the direction (cx, cy, cs): dict(Ldot, hx, hy, hs)."""
    xn, yn, sc, B, d, ws = _nce_args(xn, yn, scale)
    cx, cy = _req(cx, "cx"), _req(cy, "cy")
    if cx.shape != xn.shape or cy.shape != yn.shape:
        raise ValueError("direction shapes must match the features")
    csv = _scalar(cs, xn.device, "cs")
    Ldot, hs = torch.empty(1, device=xn.device), torch.empty(1, device=xn.device)
    hx, hy = torch.empty_like(xn), torch.empty_like(yn)
    check(lib().vldd_infonce_hvp(_ptr(xn), _ptr(yn), _ptr(sc), _ptr(cx), _ptr(cy), _ptr(csv), B, d, _ptr(Ldot), _ptr(hx),
                                 _ptr(hy), _ptr(hs), _ptr(ws), ws.numel(), _stream()), "infonce_hvp")
    return dict(Ldot=Ldot[0], hx=hx, hy=hy, hs=hs[0])


class UnrollWorkspace:
    """Caller-owned device state of the unroll engine, sized once per (N, B, K, dt, d).

    Besides the engine's scratch it owns the persistent argument / result buffers (perms, masks, out5, ce, dY, dU):
    the engine replays a CUDA graph keyed on the addresses it is given, so stable addresses mean one capture.
    """

    def __init__(self, N: int, B: int, K: int, dt: int, d: int, device):
        self.key = (N, B, K, dt, d)
        nbytes = lib().vldd_unrolled_match_workspace_bytes(N, B, K, dt, d)
        if nbytes == 0:
            check(-1, "unrolled_match_workspace_bytes")
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.perms = torch.empty(max(K, 1), B, dtype=torch.int64, device=device)
        self.masks = None
        # results live in ONE packed buffer [dU | dY | out5 | pad]: a multi-GPU run all-reduces it with a single collective and
        # no concatenation (out5 = {num, den, loss, dloss/dlr, dloss/dscale} is summed along: harmless, the loss sum is reported)
        n_u, n_y = N * d, N * dt
        self.pack = torch.zeros(n_u + n_y + 8, dtype=torch.float32, device=device)
        self.dU = self.pack[:n_u].view(N, d)
        self.dY = self.pack[n_u:n_u + n_y].view(N, dt)
        self.out5 = self.pack[n_u + n_y:n_u + n_y + 5]
        self.ce = torch.empty(max(K, 1), dtype=torch.float32, device=device)
        self.skipped = torch.zeros(1, dtype=torch.int32, device=device)     # raised by outer_update on a non-finite loss
        self.theta_K = None
        self.lr = torch.empty(1, dtype=torch.float32, device=device)
        self.scale = torch.empty(1, dtype=torch.float32, device=device)


def fill_dropout_masks(workspace: "UnrollWorkspace", p: float, rng_state: torch.Tensor) -> torch.Tensor:
    """Fresh pre-scaled dropout masks (0 or 1/(1-p)) for the K student steps, drawn in place in the workspace by the
    library's own Philox kernel (no torch kernel involved); advances `rng_state`.

    The reference's students run in train mode (distill.py:446-447), so every forward of text_projection draws a new
    nn.Dropout(0.1) mask (networks.py:636,643); the engine replays the SAME masks in its reverse sweep.  The unroll
    engine normally draws them itself (``unrolled_match(dropout_p=..., rng_state=...)``); this is the stand-alone form.
    """
    N, B, K, dt, d = workspace.key
    if workspace.masks is None:
        workspace.masks = torch.empty(max(K, 1), B, d, dtype=torch.float32, device=workspace.buf.device)
    return dropout_masks(None, p, rng_state, advance=True, out=workspace.masks)


def make_rng_state(seed: int, device) -> torch.Tensor:
    """{seed, draws so far} of the engine's Philox generator as two uint64 in device memory (viewed as int64 by torch)."""
    return torch.tensor([int(seed) & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=device)


def dropout_masks(shape, p: float, rng_state: torch.Tensor, advance: bool = True, out: torch.Tensor | None = None) -> torch.Tensor:
    """Pre-scaled dropout masks (0 or 1/(1-p)) from the engine's own generator (vldd_dropout_masks)."""
    st = _req(rng_state, "rng_state", torch.int64)
    out = torch.empty(shape, dtype=torch.float32, device=st.device) if out is None else _req(out, "out")
    check(lib().vldd_dropout_masks(_ptr(out), out.numel(), float(p), _ptr(st), int(bool(advance)), _stream()), "dropout_masks")
    return out


def outer_update(U, gU, bufU, lr_img: float, Y, gY, bufY, lr_txt: float, syn_lr_img, syn_lr_txt, g_lr_img, g_lr_txt, buf_lr,
                 lr_lr: float, momentum: float, first: bool, grad_scale: float = 1.0, loss=None, skipped=None) -> None:
    """The three SGD(momentum) steps of one outer iteration in one launch (distill.py:233-241, 603-613), in place.

    syn_lr_img / syn_lr_txt: one-element device tensors updated through buf_lr[0] / buf_lr[1]; g_lr_img may be None (the
    logit scale is not tied to syn_lr_img: distill_original.py:430).  A non-finite `loss` skips the update and sets `skipped`.
    """
    for t, n in ((U, "U"), (gU, "gU"), (bufU, "bufU"), (Y, "Y"), (gY, "gY"), (bufY, "bufY"), (buf_lr, "buf_lr")):
        _req(t, n)
        if not t.is_contiguous():
            raise ValueError(f"{n} must be contiguous for the in-place update")
    check(lib().vldd_outer_update(_ptr(U), _ptr(gU), _ptr(bufU), U.numel(), float(lr_img), _ptr(Y), _ptr(gY), _ptr(bufY),
                                  Y.numel(), float(lr_txt), _ptr(syn_lr_img), _ptr(syn_lr_txt), _ptr(g_lr_img), _ptr(g_lr_txt),
                                  _ptr(buf_lr), float(lr_lr), float(momentum), int(bool(first)), float(grad_scale), _ptr(loss),
                                  _ptr(skipped), _stream()), "outer_update")


def unrolled_match(theta0, theta_tgt, Y, U, lr, scale, perms, masks=None, workspace: UnrollWorkspace | None = None,
                   want_theta_K: bool = False, dropout_p: float = 0.0, rng_state: torch.Tensor | None = None,
                   clone_results: bool = True):
    """distill.py:509-606 for one expert segment.  Returns dict(out5=[num,den,loss,dlr,dscale], ce, dY, dU[, theta_K]).

    dropout_p > 0 with an `rng_state` (make_rng_state): the engine draws fresh masks itself inside its launch graph
    (train-mode students, distill.py:446-447); they can be read back from ``workspace.masks``.  `masks` given: used as is.
    clone_results=False returns views of the workspace's result buffers (valid until the next call on that workspace)
    and reads one-element device tensors `lr` / `scale` in place: no torch kernel is launched for the call.
    """
    theta0, theta_tgt = _req(_flat(theta0), "theta0"), _req(_flat(theta_tgt), "theta_tgt")
    Y, U = _req(Y, "Y"), _req(U, "U")
    perms = _req(perms, "perms", torch.int64)
    K, B = perms.shape
    N, dt = Y.shape
    d = U.shape[1]
    if U.shape[0] != N:
        raise ValueError("Y and U must have the same number of rows")
    if theta0.numel() != head_numel(dt, d) or theta_tgt.numel() != theta0.numel():
        raise ValueError(f"theta has {theta0.numel()} elements, expected {head_numel(dt, d)}")
    dev = Y.device
    if workspace is None or workspace.key != (N, B, K, dt, d):
        workspace = UnrollWorkspace(N, B, K, dt, d, dev)
    ws = workspace
    lr_t, sc_t = _scalar(lr, dev, "lr"), _scalar(scale, dev, "scale")
    if clone_results:                       # general API: private copies, so the caller may change lr / scale afterwards
        ws.lr.copy_(lr_t)
        ws.scale.copy_(sc_t)
        lr_t, sc_t = ws.lr, ws.scale
    # (the engine reaches the index array through a device-side pointer slot and keeps its own copy for the reverse sweep, so
    #  any device address may be passed from call to call without a copy here and without a new launch graph)
    perms_arg = perms if K > 0 else ws.perms
    m_ptr = None
    rng = None
    if dropout_p > 0.0 and K > 0:
        if masks is not None:
            raise ValueError("give either explicit masks or dropout_p, not both")
        if rng_state is None:
            raise ValueError("dropout_p > 0 needs an rng_state (ops.make_rng_state)")
        rng = _req(rng_state, "rng_state", torch.int64)
        if ws.masks is None:
            ws.masks = torch.empty(K, B, d, dtype=torch.float32, device=dev)
        m_ptr = ws.masks
    elif masks is not None:
        masks = _req(masks, "masks")
        if tuple(masks.shape) != (K, B, d):
            raise ValueError(f"masks must be [{K},{B},{d}]")
        if ws.masks is None:
            ws.masks = torch.empty(K, B, d, dtype=torch.float32, device=dev)
        if masks.data_ptr() != ws.masks.data_ptr():
            ws.masks.copy_(masks)
        m_ptr = ws.masks
    if want_theta_K and ws.theta_K is None:
        ws.theta_K = torch.empty_like(theta0)
    thK = ws.theta_K if want_theta_K else None
    check(lib().vldd_unrolled_match(_ptr(theta0), _ptr(theta_tgt), _ptr(Y), _ptr(U), _ptr(lr_t), _ptr(sc_t),
                                    _ptr(perms_arg), _ptr(m_ptr), float(dropout_p if rng is not None else 0.0), _ptr(rng),
                                    N, B, K, dt, d, _ptr(ws.out5), _ptr(ws.ce), _ptr(ws.dY), _ptr(ws.dU), _ptr(thK),
                                    _ptr(ws.buf), ws.buf.numel(), _stream()), "unrolled_match")
    if not clone_results:
        res = dict(out5=ws.out5, ce=ws.ce[:K], dY=ws.dY, dU=ws.dU, workspace=ws)
        if want_theta_K:
            res["theta_K"] = thK
        return res
    # results are copied out of the persistent buffers (1.2 MB at Flickr shape) so that they survive the next call
    res = dict(out5=ws.out5.clone(), ce=ws.ce[:K].clone(), dY=ws.dY.clone(), dU=ws.dU.clone(), workspace=ws)
    if want_theta_K:
        res["theta_K"] = thK.clone()
    return res
