"""GPU parity of the section-8f rows through the C ABI: vldd_clip_loss (CLIPModel_full.forward + backward,
networks.py:845-889 / epoch.py:59-98) and vldd_nearest_rows (nearest_neighbor, distill.py:89-95).

Floating point: loss and gradients within 1e-4 relative of the fp64 oracle and of the reference's own fp32 outputs
(tests/golden/clip_forward.npz); top-1 counters and nearest indices are integers and must match exactly."""
import os
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import distill_ref as R, retrieval_ref as RR
from test_oracle_widen import clip_case

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


@pytest.mark.parametrize("tag", ["small", "drop", "flickr"])
def test_clip_loss_matches_oracle_and_reference(tag):
    from multimodal_dataset_distillation_b200 import ops
    pr, U, mask, gold, (B, dt, d) = clip_case(tag)
    th = pr["theta0"].double().requires_grad_(True)
    Y = pr["Y"].double().requires_grad_(True)
    Ud = U.double().requires_grad_(True)
    loss, top_r, top_c = R.clip_forward_ref(th, Y, Ud, mask=None if mask is None else mask.double(), dt=dt, d=d)
    loss.backward()
    got = ops.clip_loss(pr["theta0"].cuda(), pr["Y"].cuda(), U.cuda(), ops.LOGIT_SCALE_EVAL, None if mask is None else mask.cuda())
    assert abs(float(got["loss"]) - float(loss.detach())) <= RTOL * abs(float(loss.detach()))
    assert got["top1"].tolist() == [top_r, top_c]
    assert rel(got["g_theta"], th.grad) < RTOL
    assert rel(got["dY"], Y.grad) < RTOL
    assert rel(got["dU"], Ud.grad) < RTOL
    # the reference's own fp32 run
    assert abs(float(got["loss"]) - float(gold["loss"])) <= RTOL * abs(float(gold["loss"]))
    assert got["top1"].sum().item() / 2 == float(gold["acc"])
    small = d <= 64
    assert rel(got["dY"], gold["dY"]) < 2 * RTOL
    assert rel(got["dU"] if small else got["dU"][::7], gold["dU"]) < 2 * RTOL
    assert rel(got["g_theta"] if small else got["g_theta"][::997], gold["g_theta"]) < 2 * RTOL


def test_top1_counts_first_index_on_ties():
    """torch.argmax picks the first maximum: duplicated captions / images only credit the lower index."""
    from multimodal_dataset_distillation_b200 import ops
    B, dt, d = 8, 10, 16
    pr = R.make_problem(N=B, B=B, K=1, dt=dt, d=d, seed=2)
    Y, U = pr["Y"].clone(), pr["U"].clone()
    Y[5] = Y[2]          # captions 2 and 5 identical -> columns 2 and 5 of the logits identical
    U[6] = U[1]          # images 1 and 6 identical   -> rows 1 and 6 identical
    _, top_r, top_c = R.clip_forward_ref(pr["theta0"].double(), Y.double(), U.double(), dt=dt, d=d)
    got = ops.clip_loss(pr["theta0"].cuda(), Y.cuda(), U.cuda())
    assert got["top1"].tolist() == [top_r, top_c]


@pytest.mark.parametrize("tag", ["small", "drop"])
def test_clip_model_module_trains_like_the_reference(tag):
    """CLIPModel_full mirror: forward returns (loss, acc); loss.backward() fills the head parameters' and the image
    encoder's gradients with what the reference's autograd produces."""
    from multimodal_dataset_distillation_b200 import networks
    pr, U, mask, gold, (B, dt, d) = clip_case(tag)

    class Enc(torch.nn.Module):                       # identity "image encoder" with one parameter so grads can be checked
        def __init__(self):
            super().__init__()
            self.gain = torch.nn.Parameter(torch.ones(()))

        def forward(self, x):
            return x * self.gain

    net = networks.CLIPModel_full(types.SimpleNamespace(distill=True), image_encoder=Enc(), image_embedding=d, text_embedding=dt).cuda()
    with torch.no_grad():
        off = 0
        for p in net.text_projection.parameters():
            p.copy_(pr["theta0"][off:off + p.numel()].reshape(p.shape))
            off += p.numel()
    if mask is None:
        net.eval()
    else:
        net.train()
        net.text_projection.dropout_mask = lambda rows, device: mask.cuda()        # inject the golden's mask
    Uc = U.cuda().requires_grad_(True)
    loss, acc = net(Uc, pr["Y"].cuda(), 0)
    loss.backward()
    assert abs(float(loss.detach()) - float(gold["loss"])) <= RTOL * abs(float(gold["loss"]))
    assert acc == float(gold["acc"])
    g_theta = torch.cat([p.grad.reshape(-1) for p in net.text_projection.parameters()])
    assert rel(g_theta, gold["g_theta"]) < 2 * RTOL
    assert rel(Uc.grad, gold["dU"]) < 2 * RTOL
    assert abs(float(net.image_encoder.gain.grad) - float((torch.from_numpy(gold["dU"]) * U).sum())) <= 1e-3 * float(
        (torch.from_numpy(gold["dU"]) * U).abs().sum())


def test_clip_model_rejects_cpu_tensors():
    from multimodal_dataset_distillation_b200 import ops
    pr = R.make_problem(N=4, B=4, K=1, dt=8, d=8, seed=0)
    with pytest.raises(RuntimeError):
        ops.clip_loss(pr["theta0"], pr["Y"], pr["U"])


@pytest.mark.parametrize("tag", ["small", "mid"])
def test_nearest_rows_matches_reference(tag):
    from multimodal_dataset_distillation_b200 import ops, distill
    query, bank = RR.nearest_problem(tag)
    gold = np.load(os.path.join(GOLDEN_DIR, "nearest.npz"))[f"{tag}_idx"]
    idx, cos = ops.nearest_rows(torch.from_numpy(query).cuda(), torch.from_numpy(bank).cuda(), return_cos=True)
    assert np.array_equal(idx.cpu().numpy(), gold)
    ref_cos = np.take_along_axis(RR.cosine_matrix(query, bank), gold[:, None].astype(np.int64), axis=1)[:, 0]
    # fp32 tolerance of the north star (1e-4 relative).  Measured here: -1.2e-5, one-sided -- the tensor core's fp32
    # accumulator truncates, and a query that nearly equals its bank row makes every product positive
    assert np.allclose(cos.cpu().numpy(), ref_cos, rtol=1e-4, atol=0)
    sentences = [f"caption {i}" for i in range(bank.shape[0])]
    assert distill.nearest_neighbor(sentences, torch.from_numpy(query), bank) == [sentences[i] for i in gold]


def test_nearest_rows_edge_cases():
    from multimodal_dataset_distillation_b200 import ops
    bank = torch.randn(5, 8, device="cuda")
    assert ops.nearest_rows(torch.empty(0, 8, device="cuda"), bank).numel() == 0
    one = ops.nearest_rows(bank[3:4].clone() * 3.0, bank)
    assert one.tolist() == [3]
    with pytest.raises(ValueError):
        ops.nearest_rows(torch.randn(2, 7, device="cuda"), bank)
