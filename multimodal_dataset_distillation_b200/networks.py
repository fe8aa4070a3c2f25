"""Mirror of the two pieces of the reference's networks.py that sit on the hot path's boundary.

  ProjectionHead   networks.py:625-646   same module tree (projection, gelu, fc, dropout, layer_norm), so state dicts
                                         and ReparamModule flat layouts are interchangeable with the reference's
  CLIPModel_full   networks.py:835-889   forward(image, caption, epoch) -> (loss, acc)

The encoders are NOT rebuilt here (BASELINE.json north_star: NFNet/BERT feature extraction stays outside the hot path):
`CLIPModel_full` takes the image encoder as a module and the captions as precomputed text-encoder embeddings
(`args.distill` mode of the reference, networks.py:860-861).  From the encoder outputs on -- text head, normalise,
logits, symmetric cross-entropy, top-1 counters and the whole backward -- one C-ABI call (vldd_clip_loss) does the work;
autograd sees a single node that hands `dU` back to the image encoder and the head gradient to the head's parameters.
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class ProjectionHead(nn.Module):
    """networks.py:625-646.  `forward` is the plain module (used for state-dict / ReparamModule compatibility and by
    callers that want the embedding itself); the training loss goes through `CLIPModel_full` instead."""

    def __init__(self, embedding_dim: int, projection_dim: int = 2304, dropout: float = 0.1):
        super().__init__()
        self.projection = nn.Linear(embedding_dim, projection_dim)
        self.gelu = nn.GELU()
        self.fc = nn.Linear(projection_dim, projection_dim)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(projection_dim)

    def flat_parameters(self) -> torch.Tensor:
        """Parameters in ReparamModule order (reparam_module.py:28-51), differentiable w.r.t. each of them."""
        return torch.cat([p.reshape(-1) for p in (self.projection.weight, self.projection.bias, self.fc.weight,
                                                  self.fc.bias, self.layer_norm.weight, self.layer_norm.bias)])

    def dropout_mask(self, rows: int, device) -> torch.Tensor | None:
        """Pre-scaled mask a train-mode nn.Dropout would apply (0 or 1/(1-p)); None in eval mode or for p = 0."""
        p = self.dropout.p
        if not self.training or p <= 0.0:
            return None
        d = self.fc.out_features
        return torch.empty(rows, d, device=device).bernoulli_(1.0 - p).mul_(1.0 / (1.0 - p))

    def forward(self, x: torch.Tensor, mask: torch.Tensor | None = None) -> torch.Tensor:
        d = self.fc.out_features
        if mask is None:
            mask = self.dropout_mask(x.shape[0], x.device)
        return ops.proj_head_forward(self.flat_parameters().detach(), x.float(), d, mask)


class _ClipLoss(torch.autograd.Function):
    """loss = clip_loss(theta, Y, U); the kernels produce every first-order gradient in the forward call."""

    @staticmethod
    def forward(ctx, theta, Y, U, scale, mask):
        res = ops.clip_loss(theta.detach(), Y.detach(), U.detach(), scale, mask)
        ctx.save_for_backward(res["g_theta"], res["dY"], res["dU"])
        ctx.mark_non_differentiable(res["top1"])
        return res["loss"], res["top1"]

    @staticmethod
    def backward(ctx, gout, _gtop1):
        g_theta, dY, dU = ctx.saved_tensors
        return (gout * g_theta if ctx.needs_input_grad[0] else None, gout * dY if ctx.needs_input_grad[1] else None,
                gout * dU if ctx.needs_input_grad[2] else None, None, None)


def clip_contrastive_loss(theta: torch.Tensor, text_features: torch.Tensor, image_features: torch.Tensor,
                          scale: float = ops.LOGIT_SCALE_EVAL, mask: torch.Tensor | None = None):
    """Differentiable (first order) symmetric InfoNCE of networks.py:866-889; returns (loss, top1[2] int32)."""
    return _ClipLoss.apply(theta, text_features.float(), image_features.float(), scale, mask)


class CLIPModel_full(nn.Module):
    """networks.py:835-889 with injected encoders.

    image_encoder: any module mapping an image batch to [B, image_embedding] features (the reference builds a timm NFNet);
    captions reach `forward` as text-encoder embeddings [B, text_embedding] (the reference's `distill` branch), or as
    whatever `text_encoder` maps to such embeddings when one is given.
    """

    def __init__(self, args=None, image_encoder: nn.Module | None = None, text_encoder: nn.Module | None = None,
                 image_embedding: int = 2304, text_embedding: int = 768, temperature: float = 0.07):
        super().__init__()
        if image_encoder is None:
            raise ValueError("CLIPModel_full needs an image_encoder module: the NFNet/BERT towers are outside this library")
        self.image_encoder = image_encoder
        self.text_encoder = text_encoder
        self.image_embedding, self.text_embedding = image_embedding, text_embedding
        self.text_projection = ProjectionHead(embedding_dim=text_embedding, projection_dim=image_embedding)
        self.temperature = temperature
        self.args = args
        self.distill = bool(getattr(args, "distill", True))
        self.logit_scale = float(1.0 / temperature)            # networks.py:878: exp(log(1/0.07))

    def forward(self, image, caption, epoch=None):
        image_features = self.image_encoder(image).float()
        if isinstance(caption, torch.Tensor):
            text_features = caption.to(image_features.device).float()
        elif self.text_encoder is not None:
            text_features = self.text_encoder(caption, device=image_features.device).float()
        else:
            raise TypeError("captions must be precomputed text embeddings (a tensor) when no text_encoder is attached")
        mask = self.text_projection.dropout_mask(text_features.shape[0], image_features.device)
        loss, top1 = clip_contrastive_loss(self.text_projection.flat_parameters(), text_features, image_features,
                                           self.logit_scale, mask)
        acc = top1.sum().item() / 2                               # networks.py:884-886
        return loss, acc
