"""CUDA streaming kernels vs the oracle (oracle/distill_ref.py), through the C ABI."""
import pytest
import torch

from oracle import distill_ref as R

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 3, 4, 5, 1023, 4096, 100003, 7087104]


@pytest.mark.parametrize("n", SIZES)
def test_flat_sgd_step(n):
    from multimodal_dataset_distillation_b200 import ops
    g = torch.Generator().manual_seed(n)
    th, gr = torch.randn(n, generator=g), torch.randn(n, generator=g)
    lr = torch.tensor(0.137)
    out = ops.flat_sgd_step(th.cuda(), gr.cuda(), lr.cuda())
    ref = R.flat_sgd_step_ref(th, gr, lr)
    # fp32 a - lr*b: one fma vs mul+sub -> <= 1 ulp of the product
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-6, atol=1e-6)


def test_flat_sgd_step_unaligned_and_2d():
    from multimodal_dataset_distillation_b200 import ops
    base = torch.randn(4099).cuda()
    gr = torch.randn(4099).cuda()
    out = ops.flat_sgd_step(base[1:].unsqueeze(0), gr[1:], 0.5)      # [1,P] DataParallel convention + odd alignment
    torch.testing.assert_close(out, base[1:] - 0.5 * gr[1:], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("n", [1, 5, 4096, 100003, 7087104])
def test_match_loss(n):
    from multimodal_dataset_distillation_b200 import ops
    g = torch.Generator().manual_seed(n + 1)
    th0 = torch.randn(n, generator=g)
    tgt = th0 + 0.01 * torch.randn(n, generator=g)
    thK = th0 + 0.003 * torch.randn(n, generator=g)
    out = ops.match_loss(thK.cuda(), tgt.cuda(), th0.cuda()).cpu()
    num, den = R.match_loss_ref(thK.double(), tgt.double(), th0.double())
    assert abs(out[0].item() - num.item()) <= 1e-5 * num.item() + 1e-12
    assert abs(out[1].item() - den.item()) <= 1e-5 * den.item() + 1e-12
    assert abs(out[2].item() - (num / den).item()) <= 1e-5 * (num / den).item()
    # determinism: bit-identical on repeat
    out2 = ops.match_loss(thK.cuda(), tgt.cuda(), th0.cuda()).cpu()
    assert torch.equal(out, out2)
    a = ops.match_loss_bwd(thK.cuda(), tgt.cuda(), out.cuda()).cpu()
    ref = 2 * (thK - tgt) / out[1]
    torch.testing.assert_close(a, ref, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("n", [1, 7, 30720, 15052800])
def test_momentum_sgd(n):
    from multimodal_dataset_distillation_b200 import ops
    g = torch.Generator().manual_seed(3)
    p = torch.randn(n, generator=g)
    buf = torch.zeros(n)
    pc, bc = p.clone().cuda(), buf.clone().cuda()
    for it in range(3):
        gr = torch.randn(n, generator=g)
        p, buf = R.momentum_sgd_ref(p, gr, buf, 1000.0, 0.5, it == 0)
        ops.momentum_sgd_(pc, gr.cuda(), bc, 1000.0, 0.5, it == 0)
    torch.testing.assert_close(pc.cpu(), p, rtol=1e-5, atol=1e-3)
    torch.testing.assert_close(bc.cpu(), buf, rtol=1e-6, atol=1e-6)


def test_momentum_sgd_matches_torch_optim():
    from multimodal_dataset_distillation_b200 import ops
    p = torch.randn(1000, requires_grad=True)
    opt = torch.optim.SGD([p], lr=10.0, momentum=0.5)        # distill.py:233
    pc, bc = p.detach().clone().cuda(), torch.zeros(1000).cuda()
    for it in range(4):
        gr = torch.randn(1000)
        p.grad = gr.clone()
        opt.step()
        ops.momentum_sgd_(pc, gr.cuda(), bc, 10.0, 0.5, it == 0)
    torch.testing.assert_close(pc.cpu(), p.detach(), rtol=1e-5, atol=1e-5)


def test_cpu_tensor_is_an_error():
    from multimodal_dataset_distillation_b200 import ops
    with pytest.raises(RuntimeError):
        ops.flat_sgd_step(torch.randn(8), torch.randn(8), 0.1)
