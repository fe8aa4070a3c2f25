"""The bidirectional InfoNCE loss of distill.py:548-551 as an autograd node PyTorch can differentiate twice.

"Mode B" of SURVEY.md section 8a/D1: when the image tower is a real network under PyTorch autograd (the reference runs
pixels through an NFNet wrapped in ReparamModule, distill.py:524-545), the unroll is driven by
``torch.autograd.grad(loss, params, create_graph=True)`` (distill.py:562-567) and ``grand_loss.backward()`` (606) then
differentiates those gradients again.  `infonce_loss` gives that machinery a loss node whose first derivative
(`vldd_infonce_grad`) and whose Hessian-vector product (`vldd_infonce_hvp`) are CUDA kernels, so the contrastive part of
the double backward is two C-ABI calls instead of the ~40 decomposed softmax / log / matmul backward kernels autograd
would record.  The text tower can be any module as well; the engine of `distill.UnrolledMatch` ("Mode A") remains the
fast path when the image side is a frozen embedding.
"""
from __future__ import annotations

import torch
from torch.autograd.function import once_differentiable

from . import ops


class _InfoNCEGrad(torch.autograd.Function):
    """(xn, yn, scale, gout) -> gout * grad L.  Its backward is the Hessian-vector product."""

    @staticmethod
    def forward(ctx, xn, yn, scale, gout):
        res = ops.infonce_grad(xn, yn, scale)
        ctx.save_for_backward(xn, yn, scale, gout)
        return gout * res["dxn"], gout * res["dyn"], (gout * res["dscale"]).reshape(scale.shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, cx, cy, cs):
        xn, yn, scale, gout = ctx.saved_tensors
        zero = lambda t, ref: torch.zeros_like(ref) if t is None else t
        h = ops.infonce_hvp(xn, yn, scale, zero(cx, xn).contiguous(), zero(cy, yn).contiguous(), zero(cs, scale))
        return gout * h["hx"], gout * h["hy"], (gout * h["hs"]).reshape(scale.shape), h["Ldot"].reshape(gout.shape)


class InfoNCE(torch.autograd.Function):
    """loss = (CE(S) + CE(S^T)) / 2, S = scale * xn yn^T, for row-normalised xn, yn [B, d] and a 0-dim `scale` tensor."""

    @staticmethod
    def forward(ctx, xn, yn, scale):
        ctx.save_for_backward(xn, yn, scale)
        return ops.infonce_grad(xn.detach(), yn.detach(), scale.detach())["loss"]

    @staticmethod
    def backward(ctx, gout):
        xn, yn, scale = ctx.saved_tensors
        return _InfoNCEGrad.apply(xn, yn, scale, gout)


def infonce_loss(image_features: torch.Tensor, text_features: torch.Tensor, scale) -> torch.Tensor:
    """distill.py:533,546-551: row-normalise both sides (no epsilon), logits = scale * X Y^T, symmetric cross-entropy.

    `scale` may be a Python number or a 0-dim tensor that requires grad (the fork uses the learnable syn_lr_img as the
    logit scale, distill.py:548).  Twice differentiable in all three arguments.
    """
    x = image_features.float()
    y = text_features.float()
    xn = (x / x.norm(dim=1, keepdim=True)).contiguous()
    yn = (y / y.norm(dim=1, keepdim=True)).contiguous()
    if not isinstance(scale, torch.Tensor):
        scale = torch.tensor(float(scale), device=x.device)
    return InfoNCE.apply(xn, yn, scale.float())
