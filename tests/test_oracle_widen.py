"""Oracle restatements of the section-8f rows against goldens made by executing the reference's own source
(tests/golden/make_golden_widen.py): CLIPModel_full.forward (networks.py:845-889) and nearest_neighbor (distill.py:89-95)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import distill_ref as R, retrieval_ref as RR

CLIP_CASES = {"small": (12, 16, 24, False), "drop": (20, 24, 40, True), "flickr": (100, 768, 2304, True)}


def clip_case(tag):
    B, dt, d, drop = CLIP_CASES[tag]
    pr = R.make_problem(N=B, B=B, K=1, dt=dt, d=d, seed=21, dropout=drop)
    z = np.load(os.path.join(GOLDEN_DIR, "clip_forward.npz"))
    gold = {k[len(tag) + 1:]: z[k] for k in z.files if k.startswith(tag + "_")}
    return pr, torch.from_numpy(gold["U"]), (pr["masks"][0] if drop else None), gold, (B, dt, d)


@pytest.mark.parametrize("tag", ["small", "drop", "flickr"])
def test_clip_forward_oracle_matches_reference_forward(tag):
    pr, U, mask, gold, (B, dt, d) = clip_case(tag)
    th = pr["theta0"].double().requires_grad_(True)
    Y = pr["Y"].double().requires_grad_(True)
    Ud = U.double().requires_grad_(True)
    loss, top_r, top_c = R.clip_forward_ref(th, Y, Ud, mask=None if mask is None else mask.double(), dt=dt, d=d)
    loss.backward()
    # the reference ran in fp32 (it casts with .float()); the fp64 oracle must agree to fp32 rounding
    assert abs(float(loss.detach()) - float(gold["loss"])) <= 2e-5 * abs(float(gold["loss"]))
    assert (top_r + top_c) / 2 == float(gold["acc"])
    small = d <= 64
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    assert rel(Y.grad.numpy(), gold["dY"]) < 2e-4
    assert rel(Ud.grad.numpy() if small else Ud.grad.numpy()[::7], gold["dU"]) < 2e-4
    assert rel(th.grad.numpy() if small else th.grad.numpy()[::997], gold["g_theta"]) < 2e-4


@pytest.mark.parametrize("tag", ["small", "mid"])
def test_nearest_neighbor_oracle_matches_reference(tag):
    query, bank = RR.nearest_problem(tag)
    gold = np.load(os.path.join(GOLDEN_DIR, "nearest.npz"))[f"{tag}_idx"]
    got = RR.nearest_neighbor_ref(query, bank)
    assert np.array_equal(got, gold)
    if tag == "small":
        assert got[0] == 3            # duplicate rows 3 and 10: the first index wins
