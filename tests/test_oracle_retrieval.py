"""CPU: the retrieval oracle against outputs of the REFERENCE's own itm_eval / epoch_test (tests/golden)."""
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from conftest import GOLDEN_DIR
from oracle import retrieval_ref as RR


def _case_inputs(case):
    img, txt = RR.synthetic_retrieval(case["I"], case["C"], case["D"], seed=case["seed"])
    S = (np.float32(14.285714) * img) @ txt.T
    if case["quant"]:
        S = np.round(S * case["quant"]) / np.float32(case["quant"])
    S = S.astype(np.float32)
    St = np.ascontiguousarray(S.T)
    if case["fill"]:
        S, St = RR.topk_fill_ref(S, 128), RR.topk_fill_ref(St, 128)
    return S, St, RR.flickr_maps(case["I"], case["C"])


def test_oracle_matches_reference_itm_eval(golden):
    for case in golden["retrieval"]:
        if case["I"] >= 1000 and case["fill"]:
            continue                                    # same ranks as the unfilled case; skip the slow top-k fill on CPU
        S, St, (txt2img, img2txt) = _case_inputs(case)
        assert abs(float(np.float64(S).sum()) - case["score_checksum"]) < 1e-6 * max(1.0, abs(case["score_checksum"]))
        r_i, r_t = RR.ranks_i2t(S, img2txt), RR.ranks_t2i(St, txt2img)
        res = RR.recall_dict(r_i, r_t)
        assert tuple(res.keys()) == RR.RESULT_KEYS
        if case["quant"]:
            lo_i, hi_i = RR.rank_bounds(S, [img2txt[i] for i in range(case["I"])])
            lo_t, hi_t = RR.rank_bounds(St, [[txt2img[t]] for t in range(St.shape[0])])
            best, worst = RR.recall_dict(lo_i, lo_t), RR.recall_dict(hi_i, hi_t)
            assert (lo_i <= r_i).all() and (r_i <= hi_i).all() and (lo_t <= r_t).all() and (r_t <= hi_t).all()
            for ref in (case["fork"], case["orig"]):
                for k, v in ref.items():
                    assert worst[k] - 1e-9 <= v <= best[k] + 1e-9, (case["name"], k)
        else:
            for ref in (case["fork"], case["orig"]):
                for k, v in ref.items():
                    assert res[k] == pytest.approx(v, abs=1e-12), (case["name"], k)
        if "ranks_i2t" in case:
            assert np.array_equal(r_i, np.asarray(case["ranks_i2t"], dtype=np.int32))
            assert np.array_equal(r_t, np.asarray(case["ranks_t2i"], dtype=np.int32))


def test_oracle_matches_reference_epoch_test():
    """epoch_original.epoch_test run in the build container with a fake model -> epoch_test_small.npz."""
    import torch
    from oracle import distill_ref as DR
    z = np.load(os.path.join(GOLDEN_DIR, "epoch_test_small.npz"))
    I, C, D, dt = (int(x) for x in z["dims"])
    txt = DR.head_forward(torch.from_numpy(z["theta"]), torch.from_numpy(z["bert"]), dt, D).numpy()
    txt = RR.l2_normalise(txt)
    img = RR.l2_normalise(RR.l2_normalise(z["feats"]))
    s1, s2 = RR.epoch_test_ref(img, txt)
    for got, ref in ((s1, z["s_i2t"]), (s2, z["s_t2i"])):
        kept_g, kept_r = got > -100, ref > -100
        assert (kept_g.sum(axis=1) == 128).all() and (kept_r.sum(axis=1) == 128).all()
        assert (kept_g != kept_r).sum() <= 4
        both = kept_g & kept_r
        np.testing.assert_allclose(got[both], ref[both], rtol=2e-5, atol=2e-5)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 12), st.integers(1, 40), st.integers(0, 2**31 - 1), st.integers(1, 6))
def test_rank_definition_is_a_stable_descending_sort(rows, cols, seed, levels):
    rng = np.random.default_rng(seed)
    S = rng.integers(0, levels, size=(rows, cols)).astype(np.float32)
    for r in range(rows):
        order = np.argsort(-S[r], kind="stable")
        for c in range(cols):
            assert RR.stable_desc_rank(S[r], c) == int(np.where(order == c)[0][0])


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 10), st.integers(1, 4), st.integers(0, 2**31 - 1))
def test_csr_and_dict_forms_agree(n_img, caps, seed):
    rng = np.random.default_rng(seed)
    T = n_img * caps
    S = rng.standard_normal((n_img, T)).astype(np.float32)
    txt2img, img2txt = RR.flickr_maps(n_img, caps)
    ptr = np.arange(0, T + 1, caps, dtype=np.int32)
    idx = np.arange(T, dtype=np.int32)
    assert np.array_equal(RR.ranks_i2t(S, img2txt), RR.ranks_vectorised(S, ptr, idx))


def test_topk_fill_keeps_exactly_k_and_is_idempotent():
    rng = np.random.default_rng(0)
    S = rng.standard_normal((6, 300)).astype(np.float32)
    F = RR.topk_fill_ref(S, 128)
    assert ((F > -100).sum(axis=1) == 128).all()
    assert np.array_equal(RR.topk_fill_ref(F, 128), F)
    # recall@<=10 is invariant under the fill
    txt2img, img2txt = RR.flickr_maps(6, 50)
    a = RR.ranks_i2t(S, img2txt)
    b = RR.ranks_i2t(F, img2txt)
    assert np.array_equal(a < 10, b < 10) and np.array_equal(a[a < 128], b[a < 128])


def test_c_oracle_agrees_with_numpy_oracle_and_reference_goldens(golden):
    """oracle/c/itm_eval_ref.c (plain C, built into oracle/_build/) against the numpy restatement on every golden case, and
    against the reference's own itm_eval numbers on the tie-free ones."""
    for case in golden["retrieval"]:
        if case["I"] >= 1000 and case["fill"]:
            continue
        S, St, (txt2img, img2txt) = _case_inputs(case)
        res, r_i, r_t = RR.itm_eval_c(S, St, txt2img, img2txt, return_ranks=True)
        assert np.array_equal(r_i, RR.ranks_i2t(S, img2txt)) and np.array_equal(r_t, RR.ranks_t2i(St, txt2img))
        ref = RR.recall_dict(r_i, r_t)
        assert tuple(res.keys()) == RR.RESULT_KEYS and all(abs(res[k] - ref[k]) < 1e-12 for k in res)
        if not case["quant"] and not case["fill"]:
            for k, v in case["fork"].items():
                assert abs(res[k] - v) < 1e-9, (case["name"], k)


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 12), st.integers(1, 4), st.integers(0, 2 ** 31 - 1), st.sampled_from([0, 2, 8]))
def test_c_oracle_random_ties(n_img, caps, seed, quant):
    rng = np.random.default_rng(seed)
    S = rng.standard_normal((n_img, n_img * caps)).astype(np.float32)
    if quant:
        S = (np.round(S * quant) / quant).astype(np.float32)          # heavy ties: index tie-break must agree
    St = np.ascontiguousarray(S.T)
    txt2img, img2txt = RR.flickr_maps(n_img, caps)
    res, r_i, r_t = RR.itm_eval_c(S, St, txt2img, img2txt, return_ranks=True)
    assert np.array_equal(r_i, RR.ranks_i2t(S, img2txt)) and np.array_equal(r_t, RR.ranks_t2i(St, txt2img))
