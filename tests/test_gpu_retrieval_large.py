"""Retrieval parity at BASELINE.json's full sizes (configs[3], configs[4]) and on the reference's real Flickr30K maps.

The rank arithmetic is integer: given the same fp32 score matrix the ranks must equal the oracle's bit for bit.  The
score matrix itself is floating point (3xTF32 on the tensor cores vs numpy sgemm): compared within 1e-4 relative
(norm-wise: max|got - ref| <= 1e-4 * max|ref|), the tolerance BASELINE.json's north_star states for fp32.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import retrieval_ref as RR

pytestmark = pytest.mark.gpu
SCALE = 14.285714


def _device_set(I, C, D, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    img = torch.randn(I, D, generator=g, device="cuda")
    txt = torch.randn(I * C, D, generator=g, device="cuda") + 0.15 * img.repeat_interleave(C, dim=0)
    img = (img / img.norm(dim=1, keepdim=True)).contiguous()
    txt = (txt / txt.norm(dim=1, keepdim=True)).contiguous()
    T = I * C
    t2i = (torch.arange(T, device="cuda", dtype=torch.int32) // C).contiguous()
    ptr = (torch.arange(I + 1, device="cuda", dtype=torch.int32) * C).contiguous()
    idx = torch.arange(T, device="cuda", dtype=torch.int32)
    return img, txt, t2i, ptr, idx


@pytest.mark.parametrize("I,C,D", [(5000, 5, 768), (1000, 5, 2304)])
def test_fused_equals_materialised_equals_oracle_full_size(I, C, D):
    """configs[3] eval shape 5000 x 25000 (D = 768) and the Flickr test shape at the reference's true joint dim 2304
    (networks.py:835-838): vldd_sim_rank_fused == vldd_sim_rank == oracle ranking of the GPU's own score matrix."""
    from multimodal_dataset_distillation_b200 import ops
    T = I * C
    img, txt, t2i, ptr, idx = _device_set(I, C, D, seed=I + D)
    rf1, rf2 = ops.sim_rank_fused(img, txt, t2i, ptr, idx, SCALE)
    rm1, rm2 = ops.sim_rank(img, txt, t2i, ptr, idx, SCALE)
    assert torch.equal(rf1, rm1) and torch.equal(rf2, rm2)
    s1, _ = ops.sim_scores(img, txt, SCALE, want_t2i=False)
    S = s1.cpu().numpy()
    ptr_h, idx_h, t2i_h = ptr.cpu().numpy(), idx.cpu().numpy(), t2i.cpu().numpy()
    ref_i = RR.ranks_vectorised(S, ptr_h, idx_h)
    ref_t = RR.ranks_vectorised(np.ascontiguousarray(S.T), np.arange(T + 1, dtype=np.int32), t2i_h)
    assert np.array_equal(rf1.cpu().numpy(), ref_i) and np.array_equal(rf2.cpu().numpy(), ref_t)
    # the scores against numpy's fp32 GEMM on a row sample (the full product is 190 GFLOP on the host)
    rows = np.random.default_rng(0).choice(I, 64, replace=False)
    S_np = (np.float32(SCALE) * img[rows].cpu().numpy()) @ txt.cpu().numpy().T
    assert np.abs(S[rows] - S_np).max() <= 1e-4 * np.abs(S_np).max()
    # recall dict of both paths through the product's own reduction
    got = {k: v for k, v in zip(("r1", "r5", "r10"), ops.recall_counts(rf1).cpu().tolist())}
    assert got == {k: int((ref_i < n).sum()) for k, n in (("r1", 1), ("r5", 5), ("r10", 10))}


def test_fused_25k_x_125k_sampled_against_oracle():
    """configs[4] largest sweep point: 25 000 images x 125 000 captions (3.1e9 pairs; the score matrix would be 12.5 GB and is
    never formed).  Sampled image rows and caption columns are re-scored by separate small GEMMs and ranked by the oracle."""
    from multimodal_dataset_distillation_b200 import ops
    I, C, D = 25000, 5, 768
    T = I * C
    img, txt, t2i, ptr, idx = _device_set(I, C, D, seed=7)
    r1, r2 = ops.sim_rank_fused(img, txt, t2i, ptr, idx, SCALE)
    r1, r2 = r1.cpu().numpy(), r2.cpu().numpy()
    rng = np.random.default_rng(1)
    rows = np.sort(rng.choice(I, 48, replace=False))
    cols = np.sort(rng.choice(T, 48, replace=False))
    S_rows = ops.sim_scores(img[torch.from_numpy(rows).cuda()].contiguous(), txt, SCALE, want_t2i=False)[0].cpu().numpy()   # [48, T]
    S_cols = ops.sim_scores(img, txt[torch.from_numpy(cols).cuda()].contiguous(), SCALE, want_t2i=False)[0].cpu().numpy()   # [I, 48]
    img2txt = {k: list(range(C * int(i), C * int(i) + C)) for k, i in enumerate(rows)}
    ref_rows = RR.ranks_i2t(S_rows, img2txt)
    ref_cols = RR.ranks_t2i(np.ascontiguousarray(S_cols.T), {k: int(c) // C for k, c in enumerate(cols)})
    assert np.array_equal(r1[rows], ref_rows)
    assert np.array_equal(r2[cols], ref_cols)
    # sanity of the whole result: every rank is a valid position, and the planted signal is found
    assert r1.min() >= 0 and r1.max() < T and r2.min() >= 0 and r2.max() < I
    assert (r1 < 10).mean() > 0.5


def test_itm_eval_on_the_reference_flickr30k_maps():
    """The reference's own retrieval fixture Flickr30k/ann_file/flickr30k_test.json (flickr30k_dataset.py:110-118): its
    txt2img / img2txt maps (tests/golden/flickr30k_maps.json, made by make_golden_maps.py) through itm_eval and the
    fused path at the real test-set shape 1000 x 5000."""
    from multimodal_dataset_distillation_b200 import ops, epoch
    with open(os.path.join(GOLDEN_DIR, "flickr30k_maps.json")) as f:
        maps = json.load(f)["test"]
    I, T = maps["n_img"], maps["n_txt"]
    txt2img = {t: g for t, g in enumerate(maps["txt2img"])}
    img2txt = {i: l for i, l in enumerate(maps["img2txt"])}
    assert (I, T) == (1000, 5000)
    img, txt = RR.synthetic_retrieval(I, 5, 768, seed=3)
    S = ((np.float32(SCALE) * img) @ txt.T).astype(np.float32)
    St = np.ascontiguousarray(S.T)
    got = epoch.itm_eval(S, St, txt2img, img2txt)
    want = RR.itm_eval_ref(S, St, txt2img, img2txt)
    assert got == want
    t2i, ptr, idx = ops.maps_to_arrays(txt2img, img2txt, I, T)
    dev = lambda a: torch.from_numpy(a).cuda()
    r1, r2 = ops.sim_rank_fused(dev(img), dev(txt), dev(t2i), dev(ptr), dev(idx), SCALE)
    s1, _ = ops.sim_scores(dev(img), dev(txt), SCALE, want_t2i=False)
    Sg = s1.cpu().numpy()
    assert np.array_equal(r1.cpu().numpy(), RR.ranks_i2t(Sg, img2txt))
    assert np.array_equal(r2.cpu().numpy(), RR.ranks_t2i(np.ascontiguousarray(Sg.T), txt2img))


SCREEN_CHILD = r"""
import sys, numpy as np, torch
sys.path.insert(0, {root!r})
from oracle import retrieval_ref as RR
from multimodal_dataset_distillation_b200 import ops
I, C, D, quant = {I}, {C}, {D}, {quant}
img, txt = RR.synthetic_retrieval(I, C, D, seed=11)
if quant:                                    # exact ties by the thousand, also between ground-truth captions
    img, txt = (np.round(img * 8) / 8).astype(np.float32), (np.round(txt * 8) / 8).astype(np.float32)
T = I * C
txt2img, img2txt = RR.flickr_maps(I, C)
t2i, ptr, idx = ops.maps_to_arrays(txt2img, img2txt, I, T)
dev = lambda a: torch.from_numpy(a).cuda()
f1, f2 = ops.sim_rank_fused(dev(img), dev(txt), dev(t2i), dev(ptr), dev(idx), 14.285714)
m1, m2 = ops.sim_rank(dev(img), dev(txt), dev(t2i), dev(ptr), dev(idx), 14.285714)
S = ops.sim_scores(dev(img), dev(txt), 14.285714, want_t2i=False)[0].cpu().numpy()
ok = bool(torch.equal(f1, m1) and torch.equal(f2, m2))
ok = ok and np.array_equal(f1.cpu().numpy(), RR.ranks_i2t(S, img2txt)) and np.array_equal(f2.cpu().numpy(), RR.ranks_t2i(np.ascontiguousarray(S.T), txt2img))
print("RESULT", ok, int(f1.sum()), int(f2.sum()))
sys.exit(0 if ok else 1)
"""


@pytest.mark.parametrize("I,C,D,quant,env", [
    (300, 5, 64, False, {"VLDD_SCREEN_MIN_PAIRS": "0"}),                                  # screen + decide on a small problem
    (300, 5, 64, True, {"VLDD_SCREEN_MIN_PAIRS": "0"}),                                   # tie-heavy: thousands of borderline pairs
    (300, 5, 64, True, {"VLDD_SCREEN_MIN_PAIRS": "0", "VLDD_SCREEN_CAP": "128"}),         # list overflow -> exact fall-back pass
    (257, 3, 200, False, {"VLDD_SCREEN_MIN_PAIRS": "0"}),                                 # ragged tiles, K not a multiple of 64
    (300, 5, 64, True, {"VLDD_SCREEN_MIN_PAIRS": "1000000000"}),                          # screen off: the exact count pass
])
def test_screened_count_is_bit_identical_to_materialised(I, C, D, quant, env):
    """vldd_sim_rank_fused's bf16x3 screen + exact 3xTF32 decision (and its overflow fall-back) give the ranks of the
    materialised 3xTF32 score matrix bit for bit -- on tie-free data, on heavily tied data, and when the borderline list
    overflows.  The switches are read once per process, hence the child process."""
    import subprocess
    import sys
    from conftest import ROOT
    r = subprocess.run([sys.executable, "-c", SCREEN_CHILD.format(root=ROOT, I=I, C=C, D=D, quant=quant)],
                       env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])


@pytest.mark.parametrize("I,C,D,shards,quant", [(600, 5, 768, 2, False), (600, 5, 768, 3, True), (257, 3, 64, 4, True)])
def test_caption_sharded_fused_ranking_equals_single_call(I, C, D, shards, quant):
    """Multi-GPU retrieval without the GPUs: the caption axis cut into `shards` pieces, each through the shard form of the fused
    ranking (ops.FusedRankShard: candidates -> dist.merge_candidates -> counts, summed) == one vldd_sim_rank_fused call over
    everything == the oracle on the materialised scores.  (The collectives themselves are covered by tests/test_dist_gloo.py
    and by bench.py at N > 1, which asserts the same equality across real ranks.)"""
    from multimodal_dataset_distillation_b200 import ops, dist as Dm
    img, txt = RR.synthetic_retrieval(I, C, D, seed=5)
    if quant:
        img, txt = (np.round(img * 8) / 8).astype(np.float32), (np.round(txt * 8) / 8).astype(np.float32)
    T = I * C
    txt2img, img2txt = RR.flickr_maps(I, C)
    del img2txt[3]
    img2txt[3] = []                                           # an image without any caption: rank T by convention
    t2i, ptr, idx = ops.maps_to_arrays(txt2img, img2txt, I, T)
    dev = lambda a: torch.from_numpy(a).cuda()
    img_d, txt_d, t2i_d, ptr_d, idx_d = dev(img), dev(txt), dev(t2i), dev(ptr), dev(idx)
    f1, f2 = ops.sim_rank_fused(img_d, txt_d, t2i_d, ptr_d, idx_d, SCALE)
    parts = []
    for r in range(shards):
        lo, hi = Dm.shard_bounds(T, shards, r)
        parts.append(ops.FusedRankShard(img_d, txt_d[lo:hi].contiguous(), lo, t2i_d[lo:hi].contiguous(), ptr_d, idx_d, SCALE))
    cands = [p.candidates() for p in parts]
    thr_s, thr_i = Dm.merge_candidates(torch.stack([c[0] for c in cands]), torch.stack([c[1] for c in cands]))
    outs = [p.count(thr_s, thr_i, 0) for p in parts]
    counts = sum(o[0] for o in outs)
    r1 = torch.where(thr_i >= 0, counts, torch.full_like(counts, T))
    r2 = torch.cat([o[1] for o in outs])
    assert torch.equal(r1, f1) and torch.equal(r2, f2)
    assert int(f1[3]) == T
    S = ops.sim_scores(img_d, txt_d, SCALE, want_t2i=False)[0].cpu().numpy()
    ref_maps = dict(img2txt)
    ref_maps[3] = [0]                                          # (the oracle needs some caption for image 3: ignored below)
    ref = RR.ranks_i2t(S, ref_maps)
    keep = np.arange(I) != 3
    assert np.array_equal(f1.cpu().numpy()[keep], ref[keep])
    assert np.array_equal(f2.cpu().numpy(), RR.ranks_t2i(np.ascontiguousarray(S.T), txt2img))
