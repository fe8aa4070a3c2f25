"""CPU: the distill oracle against the reference mechanism (real ReparamModule + torch double backward) goldens."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import distill_ref as R


def rel_err(got, ref):
    got, ref = torch.as_tensor(got).double(), torch.as_tensor(ref).double()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-300))


def test_flat_layout_matches_reference_reparam(golden):
    g = golden["reparam"]
    assert g["param_numel"] == R.head_numel(768, 2304) == 7087104
    assert g["names"] == ["module.projection.weight", "module.projection.bias", "module.fc.weight", "module.fc.bias",
                          "module.layer_norm.weight", "module.layer_norm.bias"]
    offs = R.head_offsets(768, 2304)
    assert [offs[k][1] for k in ("W1", "b1", "W2", "b2", "gamma", "beta")] == g["numels"]
    assert offs["W2"][0] == 1771776 and offs["gamma"][0] == 7082496


def test_head_forward_matches_reference_reparam_forward():
    z = np.load(os.path.join(GOLDEN_DIR, "reparam_small.npz"))
    out = R.head_forward(torch.from_numpy(z["theta"]).unsqueeze(0), torch.from_numpy(z["x"]), 12, 20)
    np.testing.assert_allclose(out.numpy(), z["out"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", ["small_nodrop", "small_drop", "mid_full_batch"])
def test_autograd_and_manual_oracles_match_reference_f64(golden, name):
    z = np.load(os.path.join(GOLDEN_DIR, "distill_small.npz"))
    g = golden["distill"][f"{name}_f64"]
    pr = R.make_problem(dtype=torch.float64, **g["kw"])
    for fn in (R.unrolled_match_autograd, R.unrolled_match_manual):
        res = fn(**pr)
        assert float(res.loss) == pytest.approx(g["loss"], rel=1e-12)
        assert float(res.dlr) == pytest.approx(g["dlr"], rel=1e-10)
        assert float(res.dscale) == pytest.approx(g["dscale"], rel=1e-10)
        assert [float(c) for c in res.ce] == pytest.approx(g["ce"], rel=1e-12)
        assert rel_err(res.dY, z[f"{name}_f64_dY"]) < 1e-10
        assert rel_err(res.dU, z[f"{name}_f64_dU"]) < 1e-10
        assert rel_err(res.theta_K, z[f"{name}_f64_thetaK"]) < 1e-12


def test_flickr_shape_fp32_oracle_matches_reference_fp32(golden):
    """Config 3 in fp32 through the oracle == the reference mechanism's fp32 run (same torch kernels)."""
    z = np.load(os.path.join(GOLDEN_DIR, "distill_flickr.npz"))
    g = golden["distill"]["flickr_upstream_f32"]
    pr = R.make_problem(dtype=torch.float32, **g["kw"])
    res = R.unrolled_match_manual(**pr)
    assert float(res.loss) == pytest.approx(g["loss"], rel=1e-5)
    assert float(res.dlr) == pytest.approx(g["dlr"], rel=1e-4)
    assert float(res.dscale) == pytest.approx(g["dscale"], rel=1e-4)
    assert rel_err(res.dY[::10], z["flickr_upstream_f32_dY"]) < 1e-4
    assert rel_err(res.dU[::10], z["flickr_upstream_f32_dU"]) < 1e-4


def test_streaming_refs():
    th, g = torch.randn(1000), torch.randn(1000)
    assert torch.equal(R.flat_sgd_step_ref(th, g, torch.tensor(0.1)), th - 0.1 * g)
    p = torch.randn(50, requires_grad=True)
    opt = torch.optim.SGD([p], lr=3.0, momentum=0.5)
    q, buf = p.detach().clone(), torch.zeros(50)
    for it in range(3):
        gr = torch.randn(50)
        p.grad = gr.clone()
        opt.step()
        q, buf = R.momentum_sgd_ref(q, gr, buf, 3.0, 0.5, it == 0)
    torch.testing.assert_close(q, p.detach())


def test_permutation_invariance_when_batch_is_the_whole_set():
    """B == N: InfoNCE is invariant to a joint permutation of the pairs, so the loss does not depend on perms."""
    pr = R.make_problem(N=10, B=10, K=2, dt=6, d=8, seed=4, dtype=torch.float64, lr=0.3, scale=2.0, tgt_eps=0.05)
    a = R.unrolled_match_manual(**pr)
    pr["perms"] = torch.stack([torch.arange(10), torch.arange(10)])
    b = R.unrolled_match_manual(**pr)
    assert float(a.loss) == pytest.approx(float(b.loss), rel=1e-12)
    assert rel_err(a.dY, b.dY) < 1e-10
