"""CUDA retrieval kernels vs the oracle and the reference-generated goldens, through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import retrieval_ref as RR

pytestmark = pytest.mark.gpu


def _case_inputs(case):
    img, txt = RR.synthetic_retrieval(case["I"], case["C"], case["D"], seed=case["seed"])
    S = (np.float32(14.285714) * img) @ txt.T
    if case["quant"]:
        S = np.round(S * case["quant"]) / np.float32(case["quant"])
    S = S.astype(np.float32)
    St = np.ascontiguousarray(S.T)
    if case["fill"]:
        S, St = RR.topk_fill_ref(S, 128), RR.topk_fill_ref(St, 128)
    return img, txt, S, St, RR.flickr_maps(case["I"], case["C"])


def test_itm_eval_matches_reference_goldens(golden):
    from multimodal_dataset_distillation_b200 import epoch
    for case in golden["retrieval"]:
        _, _, S, St, (txt2img, img2txt) = _case_inputs(case)
        assert abs(float(np.float64(S).sum()) - case["score_checksum"]) < 1e-6 * max(1.0, abs(case["score_checksum"]))
        res, r_i, r_t = epoch.itm_eval(S, St, txt2img, img2txt, return_ranks=True)
        if case["quant"]:
            # ties near the top: np.argsort's order among equal scores is implementation-defined (SIMD introsort), so
            # the reference's recall is only pinned to the interval spanned by the two extreme tie resolutions
            lo_i, hi_i = RR.rank_bounds(S, [img2txt[i] for i in range(case["I"])])
            lo_t, hi_t = RR.rank_bounds(St, [[txt2img[t]] for t in range(St.shape[0])])
            best, worst = RR.recall_dict(lo_i, lo_t), RR.recall_dict(hi_i, hi_t)
            for ref in (case["fork"], case["orig"]):
                for k, v in ref.items():
                    assert worst[k] - 1e-9 <= v <= best[k] + 1e-9, (case["name"], k)
                    assert worst[k] - 1e-9 <= res[k] <= best[k] + 1e-9, (case["name"], k)
        else:
            for ref in (case["fork"], case["orig"]):        # recall numbers: exact (integer counts / n)
                for k, v in ref.items():
                    assert res[k] == pytest.approx(v, abs=1e-12), (case["name"], k)
        if "ranks_i2t" in case:                 # tie-free: the reference's ranks themselves, bit-exact
            assert np.array_equal(r_i, np.asarray(case["ranks_i2t"], dtype=np.int32)), case["name"]
            assert np.array_equal(r_t, np.asarray(case["ranks_t2i"], dtype=np.int32)), case["name"]
        # ties: oracle's deterministic definition, bit-exact
        assert np.array_equal(r_i, RR.ranks_i2t(S, img2txt)), case["name"]
        assert np.array_equal(r_t, RR.ranks_t2i(St, txt2img)), case["name"]


@pytest.mark.parametrize("I,C,T_extra", [(1, 1, 0), (3, 2, 0), (17, 5, 3), (33, 1, 0), (5, 7, 2)])
def test_ranks_ragged_and_edge(I, C, T_extra):
    from multimodal_dataset_distillation_b200 import ops
    rng = np.random.default_rng(I * 100 + C)
    T = I * C + T_extra
    S = rng.integers(-3, 4, size=(I, T)).astype(np.float32)        # heavy ties
    St = rng.integers(-3, 4, size=(T, I)).astype(np.float32)
    # ragged ground truth: image i owns a random non-empty subset
    img2txt = {i: sorted(rng.choice(T, size=rng.integers(1, min(T, 6) + 1), replace=False).tolist()) for i in range(I)}
    txt2img = {t: int(rng.integers(0, I)) for t in range(T)}
    t2i, ptr, idx = ops.maps_to_arrays(txt2img, img2txt, I, T)
    r1, r2 = ops.ranks_from_scores(torch.from_numpy(S).cuda(), torch.from_numpy(St).cuda(), torch.from_numpy(t2i).cuda(),
                                   torch.from_numpy(ptr).cuda(), torch.from_numpy(idx).cuda())
    assert np.array_equal(r1.cpu().numpy(), RR.ranks_i2t(S, img2txt))
    assert np.array_equal(r2.cpu().numpy(), RR.ranks_t2i(St, txt2img))


def test_ranks_unaligned_rows():
    from multimodal_dataset_distillation_b200 import ops
    rng = np.random.default_rng(9)
    I, T = 37, 4099                      # odd row length -> scalar path; >4096 -> 256-thread kernel
    S = rng.standard_normal((I, T)).astype(np.float32)
    img2txt = {i: [int(rng.integers(0, T))] for i in range(I)}
    txt2img = {t: 0 for t in range(T)}
    t2i, ptr, idx = ops.maps_to_arrays(txt2img, img2txt, I, T)
    r1, _ = ops.ranks_from_scores(torch.from_numpy(S).cuda(), None, torch.from_numpy(t2i).cuda(),
                                  torch.from_numpy(ptr).cuda(), torch.from_numpy(idx).cuda())
    assert np.array_equal(r1.cpu().numpy(), RR.ranks_i2t(S, img2txt))


def test_recall_counts():
    from multimodal_dataset_distillation_b200 import ops
    r = torch.tensor([0, 0, 1, 4, 5, 9, 10, 11, 200, 3], dtype=torch.int32).cuda()
    assert ops.recall_counts(r).cpu().tolist() == [2, 5, 7]
    assert ops.recall_counts(torch.zeros(0, dtype=torch.int32).cuda()).cpu().tolist() == [0, 0, 0]


def test_sim_scores_and_topk_fill_vs_epoch_test_golden():
    """epoch_original.epoch_test driven in the build container -> tests/golden/epoch_test_small.npz."""
    import os
    from multimodal_dataset_distillation_b200 import ops
    from conftest import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "epoch_test_small.npz"))
    I, C, D, dt = (int(x) for x in z["dims"])
    theta = torch.from_numpy(z["theta"]).cuda()
    txt = ops.proj_head_forward(theta, torch.from_numpy(z["bert"]).cuda(), D, normalise=True)
    feats = torch.from_numpy(z["feats"]).cuda()
    img = feats / feats.norm(dim=1, keepdim=True)
    img = img / img.norm(dim=1, keepdim=True)                         # epoch_original.py:84,92 (normalised twice)
    s1, s2 = ops.sim_scores(img, txt, ops.LOGIT_SCALE_EVAL)
    f1, f2 = ops.topk_fill(s1, 128, -100.0).cpu().numpy(), ops.topk_fill(s2, 128, -100.0).cpu().numpy()
    for got, ref in ((f1, z["s_i2t"]), (f2, z["s_t2i"])):
        kept_g, kept_r = got > -100, ref > -100
        assert (kept_g.sum(axis=1) == 128).all()
        # the kept sets can differ only where the 128th/129th values are within rounding of each other
        assert (kept_g != kept_r).sum() <= 4
        both = kept_g & kept_r
        np.testing.assert_allclose(got[both], ref[both], rtol=2e-5, atol=2e-5)


def test_topk_fill_exact_vs_oracle():
    from multimodal_dataset_distillation_b200 import ops
    rng = np.random.default_rng(4)
    for rows, cols, k in [(5, 300, 128), (3, 128, 128), (4, 100, 128), (6, 1000, 7), (2, 513, 1)]:
        S = rng.standard_normal((rows, cols)).astype(np.float32)
        S[0, : cols // 2] = 0.25                                       # a big tie block crossing the threshold
        out = ops.topk_fill(torch.from_numpy(S).cuda(), k, -100.0).cpu().numpy()
        assert np.array_equal(out, RR.topk_fill_ref(S, k, -100.0)), (rows, cols, k)


def test_sim_rank_flickr_shape():
    from multimodal_dataset_distillation_b200 import ops
    img, txt = RR.synthetic_retrieval(1000, 5, 768, seed=0)
    txt2img, img2txt = RR.flickr_maps(1000, 5)
    t2i, ptr, idx = ops.maps_to_arrays(txt2img, img2txt, 1000, 5000)
    dev = lambda a: torch.from_numpy(a).cuda()
    r1, r2 = ops.sim_rank(dev(img), dev(txt), dev(t2i), dev(ptr), dev(idx), 14.285714)
    # ranks are a function of fp32 scores whose summation order differs from numpy's: compare through the
    # GPU's own score matrix (bit-exact) and against numpy scores up to near-tie flips
    s1, s2 = ops.sim_scores(dev(img), dev(txt), 14.285714)
    assert np.array_equal(r1.cpu().numpy(), RR.ranks_vectorised(s1.cpu().numpy(), ptr, idx))
    # text->image ranks are taken column-wise from the SAME matrix (no transpose is built)
    assert np.array_equal(r2.cpu().numpy(), RR.ranks_vectorised(np.ascontiguousarray(s1.cpu().numpy().T),
                                                                np.arange(5001, dtype=np.int32), t2i))
    np.testing.assert_allclose(s2.cpu().numpy(), s1.cpu().numpy().T, rtol=1e-5, atol=1e-5)
    S = (np.float32(14.285714) * img) @ txt.T
    np.testing.assert_allclose(s1.cpu().numpy(), S, rtol=1e-4, atol=1e-4)
    ref1 = RR.ranks_vectorised(S, ptr, idx)
    assert (r1.cpu().numpy() != ref1).mean() < 0.01
    res_gpu = RR.recall_dict(r1.cpu().numpy(), r2.cpu().numpy())
    res_ref = RR.recall_dict(ref1, RR.ranks_vectorised(np.ascontiguousarray(S.T), np.arange(5001, dtype=np.int32), t2i))
    for k in res_ref:
        assert abs(res_gpu[k] - res_ref[k]) <= 0.2, k


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_caption_sharded_ranks_match_single_gpu(world):
    """The per-shard kernels + the product merge rule reproduce the unsharded ranks bit for bit (shards emulated
    sequentially on one GPU; the collectives themselves are tested under gloo in tests/test_dist_gloo.py)."""
    from multimodal_dataset_distillation_b200 import dist as D, ops
    rng = np.random.default_rng(world)
    I, C, Dm = 37, 5, 32
    img, txt = RR.synthetic_retrieval(I, C, Dm, seed=world)
    img, txt = (np.round(img * 4) / 4).astype(np.float32), (np.round(txt * 4) / 4).astype(np.float32)     # ties
    T = I * C
    txt2img, img2txt = RR.flickr_maps(I, C)
    t2i, ptr, idx = ops.maps_to_arrays(txt2img, img2txt, I, T)
    dev = lambda a: torch.from_numpy(a).cuda()
    s_full, st_full = ops.sim_scores(dev(img), dev(txt), 14.285714)
    ref_i, ref_t = ops.ranks_from_scores(s_full, st_full, dev(t2i), dev(ptr), dev(idx))
    be = D.CudaBackend()
    cands, shards = [], []
    for r in range(world):
        lo, hi = D.shard_bounds(T, world, r)
        s1 = be.scores(dev(img), dev(txt[lo:hi].copy()), 14.285714)
        assert torch.equal(s1, s_full[:, lo:hi])                        # same arithmetic per element regardless of tiling
        shards.append((lo, hi, s1, None))
        cands.append(be.best_gt(s1, lo, dev(ptr), dev(idx)))
    thr_s, thr_i = D.merge_candidates(torch.stack([c[0] for c in cands]), torch.stack([c[1] for c in cands]))
    counts = sum(be.count(s1, lo, thr_s, thr_i) for lo, hi, s1, s2 in shards)
    assert torch.equal(counts, ref_i)
    got_t = torch.cat([be.ranks_t2i(s1, dev(t2i[lo:hi].copy())) for lo, hi, s1, s2 in shards])
    ref_t_cols = ops.ranks_cols(s_full, dev(t2i))
    assert torch.equal(got_t, ref_t_cols)
    assert (got_t != ref_t).float().mean() < 0.05        # ref_t ranks the separately computed transpose GEMM (rounding near ties)


@pytest.mark.parametrize("I,T,D", [(1, 1, 4), (3, 7, 8), (37, 300, 20), (300, 37, 12), (513, 1030, 64)])
def test_sim_rank_random_maps_and_ties(I, T, D):
    """Arbitrary (inconsistent) ground-truth maps, heavy ties, shapes that are not multiples of any tile."""
    from multimodal_dataset_distillation_b200 import ops
    rng = np.random.default_rng(I * 7 + T)
    img = (rng.integers(-2, 3, size=(I, D)) / 2).astype(np.float32)
    txt = (rng.integers(-2, 3, size=(T, D)) / 2).astype(np.float32)
    img2txt = {i: sorted(rng.choice(T, size=rng.integers(1, min(T, 4) + 1), replace=False).tolist()) for i in range(I)}
    txt2img = {t: int(rng.integers(0, I)) for t in range(T)}
    t2i, ptr, idx = ops.maps_to_arrays(txt2img, img2txt, I, T)
    dev = lambda a: torch.from_numpy(a).cuda()
    r1, r2 = ops.sim_rank(dev(img), dev(txt), dev(t2i), dev(ptr), dev(idx), 2.0)
    s1, _ = ops.sim_scores(dev(img), dev(txt), 2.0, want_t2i=False)
    S = s1.cpu().numpy()
    assert np.array_equal(S, (2.0 * img @ txt.T).astype(np.float32))            # small half-integers: exact in any order
    assert np.array_equal(r1.cpu().numpy(), RR.ranks_i2t(S, img2txt))
    assert np.array_equal(r2.cpu().numpy(), RR.ranks_t2i(np.ascontiguousarray(S.T), txt2img))


@pytest.mark.parametrize("I,T,D,C", [(1000, 5000, 768, 5), (130, 257, 64, 0), (513, 1030, 128, 0), (64, 64, 32, 1), (300, 37, 16, 0)])
def test_sim_rank_fused_equals_materialised(I, T, D, C):
    """GEMM with rank epilogues (no score matrix in HBM) == similarity GEMM + rank kernels, bit for bit."""
    from multimodal_dataset_distillation_b200 import ops
    rng = np.random.default_rng(I + T)
    if C:                                    # Flickr-like block structure
        img, txt = RR.synthetic_retrieval(I, T // I, D, seed=I)
        txt2img, img2txt = RR.flickr_maps(I, T // I)
    else:                                    # arbitrary, inconsistent maps + heavy ties (half-integer embeddings)
        img = (rng.integers(-2, 3, size=(I, D)) / 2).astype(np.float32)
        txt = (rng.integers(-2, 3, size=(T, D)) / 2).astype(np.float32)
        img2txt = {i: sorted(rng.choice(T, size=rng.integers(1, min(T, 4) + 1), replace=False).tolist()) for i in range(I)}
        txt2img = {t: int(rng.integers(0, I)) for t in range(T)}
    t2i, ptr, idx = ops.maps_to_arrays(txt2img, img2txt, I, T)
    dev = lambda a: torch.from_numpy(a).cuda()
    a1, a2 = ops.sim_rank(dev(img), dev(txt), dev(t2i), dev(ptr), dev(idx), 14.285714)
    b1, b2 = ops.sim_rank_fused(dev(img), dev(txt), dev(t2i), dev(ptr), dev(idx), 14.285714)
    assert torch.equal(a1, b1) and torch.equal(a2, b2)
    if not C:
        S = (np.float32(14.285714) * (img @ txt.T)).astype(np.float32)
        s1, _ = ops.sim_scores(dev(img), dev(txt), 14.285714, want_t2i=False)
        if np.array_equal(s1.cpu().numpy(), S):          # exact products: the oracle on numpy scores must agree too
            assert np.array_equal(b1.cpu().numpy(), RR.ranks_i2t(S, img2txt))
            assert np.array_equal(b2.cpu().numpy(), RR.ranks_t2i(np.ascontiguousarray(S.T), txt2img))


def test_sim_rank_fused_invalid_ground_truth():
    from multimodal_dataset_distillation_b200 import ops
    rng = np.random.default_rng(5)
    I, T, D = 40, 90, 32
    img, txt = rng.standard_normal((I, D)).astype(np.float32), rng.standard_normal((T, D)).astype(np.float32)
    t2i = rng.integers(0, I, size=T).astype(np.int32)
    t2i[3], t2i[77] = -1, I + 5                                   # captions without a valid image
    ptr = np.arange(I + 1, dtype=np.int32)
    idx = rng.integers(0, T, size=I).astype(np.int32)
    idx[7] = T + 3                                                # image whose only caption index is out of range
    dev = lambda a: torch.from_numpy(a).cuda()
    a1, a2 = ops.sim_rank(dev(img), dev(txt), dev(t2i), dev(ptr), dev(idx), 1.0)
    b1, b2 = ops.sim_rank_fused(dev(img), dev(txt), dev(t2i), dev(ptr), dev(idx), 1.0)
    assert torch.equal(a1, b1) and torch.equal(a2, b2)
    assert int(b1[7]) == T and int(b2[3]) == I and int(b2[77]) == I
