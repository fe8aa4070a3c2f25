"""Twice-differentiable InfoNCE node ("Mode B"): vldd_infonce_grad / vldd_infonce_hvp against the oracle (torch fp64 CPU
autograd of oracle/distill_ref.py::infonce), first on the raw entry points and then inside a two-tower unroll driven
exactly like distill.py:509-606 (autograd.grad(create_graph=True) per step, backward through the unroll).
Tolerance: 1e-4 relative (north star, fp32)."""
import pytest
import torch

from oracle import distill_ref as R

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def features(B, d, seed):
    g = torch.Generator().manual_seed(seed)
    x = R.row_normalise(torch.randn(B, d, generator=g, dtype=torch.float64))
    y = R.row_normalise(x + 0.7 * torch.randn(B, d, generator=g, dtype=torch.float64))
    return x, y, g


@pytest.mark.parametrize("B,d,scale", [(8, 16, 2.0), (100, 2304, 14.2857), (100, 2304, 0.1), (37, 64, 2.6593), (130, 256, 5.0)])
def test_infonce_grad_and_hvp(B, d, scale):
    from multimodal_dataset_distillation_b200 import ops
    x, y, g = features(B, d, 3)
    cx = torch.randn(B, d, generator=g, dtype=torch.float64)
    cy = torch.randn(B, d, generator=g, dtype=torch.float64)
    cs = torch.tensor(0.37, dtype=torch.float64)
    xr, yr = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    sr = torch.tensor(scale, dtype=torch.float64, requires_grad=True)
    loss = R.infonce(xr, yr, sr)
    gx, gy, gs = torch.autograd.grad(loss, (xr, yr, sr), create_graph=True)
    ldot = (gx * cx).sum() + (gy * cy).sum() + gs * cs
    hx, hy, hs = torch.autograd.grad(ldot, (xr, yr, sr))
    c = lambda t: t.float().cuda()
    got = ops.infonce_grad(c(x), c(y), scale)
    assert abs(float(got["loss"]) - float(loss)) <= RTOL * abs(float(loss))
    assert rel(got["dxn"], gx.detach()) < RTOL and rel(got["dyn"], gy.detach()) < RTOL
    assert abs(float(got["dscale"]) - float(gs)) <= RTOL * abs(float(gs)) + 1e-7
    h = ops.infonce_hvp(c(x), c(y), scale, c(cx), c(cy), float(cs))
    assert abs(float(h["Ldot"]) - float(ldot)) <= RTOL * abs(float(ldot)) + 1e-7
    assert rel(h["hx"], hx) < RTOL and rel(h["hy"], hy) < RTOL
    assert abs(float(h["hs"]) - float(hs)) <= RTOL * abs(float(hs)) + 1e-7


def _two_tower_unroll(loss_fn, img, txt, Wi, Wt, tgt_i, tgt_t, lr, scale, K):
    """distill.py:509-606 with both towers as students: x = tanh(img Wi), y = txt Wt (flat parameters)."""
    th_i, th_t = [Wi], [Wt]
    for _ in range(K):
        x = torch.tanh(img @ th_i[-1])
        y = txt @ th_t[-1]
        loss = loss_fn(x, y, scale)
        gi, gt = torch.autograd.grad(loss, (th_i[-1], th_t[-1]), create_graph=True)
        th_i.append(th_i[-1] - lr * gi)
        th_t.append(th_t[-1] - lr * gt)
    num = ((th_i[-1] - tgt_i) ** 2).sum() / ((Wi - tgt_i) ** 2).sum() + ((th_t[-1] - tgt_t) ** 2).sum() / ((Wt - tgt_t) ** 2).sum()
    return num


@pytest.mark.parametrize("B,p,q,d,K", [(12, 20, 10, 16, 2), (64, 48, 32, 128, 3)])
def test_two_tower_unroll_matches_torch_double_backward(B, p, q, d, K):
    from multimodal_dataset_distillation_b200.infonce import infonce_loss
    g = torch.Generator().manual_seed(9)
    mk = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    img, txt = mk(B, p), mk(B, q)
    Wi, Wt = mk(p, d) / p ** 0.5, mk(q, d) / q ** 0.5
    tgt_i, tgt_t = Wi + 0.05 * mk(p, d), Wt + 0.05 * mk(q, d)
    lr0, s0 = 0.3, 2.6593

    def run(dtype, dev, loss_fn):
        leaf = lambda t: t.to(dtype=dtype, device=dev).clone().requires_grad_(True)
        a, b = leaf(img), leaf(txt)
        lr, sc = leaf(torch.tensor(lr0)), leaf(torch.tensor(s0))
        cst = lambda t: t.to(dtype=dtype, device=dev)
        out = _two_tower_unroll(loss_fn, a, b, cst(Wi).requires_grad_(True), cst(Wt).requires_grad_(True), cst(tgt_i), cst(tgt_t), lr, sc, K)
        out.backward()
        return out.detach(), a.grad, b.grad, lr.grad, sc.grad

    ref = run(torch.float64, "cpu", lambda x, y, s: R.infonce(R.row_normalise(x), R.row_normalise(y), s))
    got = run(torch.float32, "cuda", infonce_loss)
    assert abs(float(got[0]) - float(ref[0])) <= RTOL * abs(float(ref[0]))
    assert rel(got[1], ref[1]) < RTOL and rel(got[2], ref[2]) < RTOL
    assert abs(float(got[3]) - float(ref[3])) <= RTOL * abs(float(ref[3])) + 1e-9
    assert abs(float(got[4]) - float(ref[4])) <= RTOL * abs(float(ref[4])) + 1e-9


def test_infonce_loss_rejects_cpu():
    from multimodal_dataset_distillation_b200.infonce import infonce_loss
    with pytest.raises(RuntimeError):
        infonce_loss(torch.randn(4, 8), torch.randn(4, 8), 1.0)
