"""Oracle for Path 1 (distill inner loop).  TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

torch-CPU restatement of the reference's unrolled student loop.  ``distill.py`` itself cannot be
imported here (top-level ``import clip`` / ``timm`` / ``kornia`` / BERT download), so the loop is
restated from the source lines cited on each function, and the flat-parameter layout is the one
``reparam_module.ReparamModule`` produces for ``networks.ProjectionHead`` (verified against the real
class by tests/golden/make_golden.py):

    theta = [projection.weight (d x dt) | projection.bias (d) | fc.weight (d x d) | fc.bias (d)
             | layer_norm.weight (d) | layer_norm.bias (d)]

Image side ("Mode A", BASELINE.json: NFNet stays outside the hot path as frozen embeddings): the
synthetic image variable is the image-encoder OUTPUT ``U [N, d]``; it has no student parameters.

Two independent implementations live here on purpose:
  * ``unrolled_match_autograd`` -- literal: autograd.grad(create_graph=True) per step, backward()
    through the unroll, exactly the reference's mechanism.
  * ``unrolled_match_manual``   -- the hand-derived forward-over-reverse sweep the CUDA engine
    implements, op for op, so a CUDA/oracle mismatch can be bisected per op.
Tests require the two to agree to ~1e-12 in float64.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

LN_EPS = 1e-5            # nn.LayerNorm default, networks.py:637
DROPOUT_P = 0.1          # networks.py:629,636


def head_numel(dt: int, d: int) -> int:
    return d * dt + d + d * d + d + d + d


def head_offsets(dt: int, d: int) -> dict:
    """Offsets of each parameter inside the flat vector (reparam_module.py:28-51 order)."""
    o, out = 0, {}
    for name, n in (("W1", d * dt), ("b1", d), ("W2", d * d), ("b2", d), ("gamma", d), ("beta", d)):
        out[name] = (o, n)
        o += n
    return out


def split_theta(theta: torch.Tensor, dt: int, d: int):
    """reparam_module.py:110-115 ``_unflatten_param``: split + view, no copies."""
    theta = torch.squeeze(theta)  # reparam_module.py:149 (DataParallel [1,P] convention)
    W1, b1, W2, b2, g, b = theta.split([d * dt, d, d * d, d, d, d])
    return W1.view(d, dt), b1, W2.view(d, d), b2, g, b


def head_forward(theta, y, dt, d, mask=None):
    """networks.py:639-646 ``ProjectionHead.forward`` with an explicit (pre-scaled) dropout mask."""
    W1, b1, W2, b2, g, b = split_theta(theta, dt, d)
    p = F.linear(y, W1, b1)
    h = F.gelu(p)                          # nn.GELU() default = exact erf form
    f = F.linear(h, W2, b2)
    if mask is not None:                   # nn.Dropout(0.1) in train mode: f * mask, mask in {0, 1/0.9}
        f = f * mask
    r = f + p
    return F.layer_norm(r, (d,), g, b, LN_EPS)


def infonce(xn, yn, scale):
    """distill.py:548-551 / distill_original.py:430-432: S = scale * X^ @ Y^T, (CE(S) + CE(S^T)) / 2."""
    logits = scale * xn @ yn.t()
    gt = torch.arange(len(logits), device=logits.device)
    return (F.cross_entropy(logits, gt) + F.cross_entropy(logits.t(), gt)) / 2


def row_normalise(x):
    """distill.py:533,546: x / x.norm(dim=1, keepdim=True) -- no epsilon."""
    return x / x.norm(dim=1, keepdim=True)


def clip_forward_ref(theta, Y, U, scale=math.exp(math.log(1 / 0.07)), mask=None, dt=768, d=2304):
    """networks.py:866-889 restated from the encoder outputs on: (loss, top1_rows, top1_cols); acc = (rows + cols) / 2.

    txt = text_projection(Y) (868-870); both sides row-normalised (873-874); logits = scale * Xn Yn^T with
    scale = exp(log(1/0.07)) (877-878); loss = (CE(logits) + CE(logits^T)) / 2 (881-882); torch.argmax over rows /
    columns compared with arange(B) (884-885).  Differentiable in theta, Y, U (loss.backward() in epoch.py:86).
    """
    xn = row_normalise(U)
    yn = row_normalise(head_forward(theta, Y, dt, d, mask))
    logits = scale * xn @ yn.t()
    gt = torch.arange(logits.shape[0])
    loss = (F.cross_entropy(logits, gt) + F.cross_entropy(logits.t(), gt)) / 2
    top_r = int((torch.argmax(logits, 1) == gt).sum())
    top_c = int((torch.argmax(logits, 0) == gt).sum())
    return loss, top_r, top_c



@dataclass
class UnrollResult:
    loss: torch.Tensor          # txt_param_loss (= grand_loss in Mode A)
    num: torch.Tensor
    den: torch.Tensor
    dY: torch.Tensor            # d loss / d text_syn      [N, dt]
    dU: torch.Tensor            # d loss / d image embeds  [N, d]
    dlr: torch.Tensor           # d loss / d syn_lr_txt
    dscale: torch.Tensor        # d loss / d logit scale (fork: add to syn_lr_img.grad, distill.py:548)
    ce: list                    # per-step contrastive loss values
    theta_K: torch.Tensor


def unrolled_match_autograd(theta0, theta_tgt, Y, U, lr, scale, perms, masks=None, dt=768, d=2304):
    """distill.py:509-606 (text tower; image side = embeddings).

    perms: LongTensor [K, B] (distill.py:510-511 ``randperm(N)[:mini_batch_size]``, injected).
    masks: optional [K, B, d] pre-scaled dropout masks (train-mode students, distill.py:446-447).
    """
    Y = Y.detach().clone().requires_grad_(True)
    U = U.detach().clone().requires_grad_(True)
    lr = lr.detach().clone().requires_grad_(True)
    scale = scale.detach().clone().requires_grad_(True)
    params = [theta0.detach().clone().requires_grad_(True)]          # distill.py:474
    ces = []
    for k in range(perms.shape[0]):
        idx = perms[k]
        x = row_normalise(U[idx])                                     # 524,533 (encoder output -> normalise)
        y = head_forward(params[-1], Y[idx], dt, d, None if masks is None else masks[k])   # 537
        y = row_normalise(y)                                          # 546
        ce = infonce(x, y, scale)                                     # 548-551
        ces.append(ce.detach())
        g = torch.autograd.grad(ce, params[-1], create_graph=True)[0]  # 565-567
        params.append(params[-1] - lr * g)                            # 583
    num = F.mse_loss(params[-1], theta_tgt, reduction="sum")          # 590
    den = F.mse_loss(theta0, theta_tgt, reduction="sum")              # 591
    loss = num / den                                                  # 597
    loss.backward()                                                   # 606
    return UnrollResult(loss.detach(), num.detach(), den.detach(), Y.grad, U.grad, lr.grad, scale.grad,
                        ces, params[-1].detach())


# ----------------------------------------------------------------------------------------------
# Hand-derived forward-over-reverse sweep (what the CUDA engine does).  Every helper below maps to
# one CUDA kernel / epilogue; names match multimodal_dataset_distillation_b200/csrc/.
# ----------------------------------------------------------------------------------------------
_SQRT1_2 = 1.0 / math.sqrt(2.0)
_INV_SQRT_2PI = 1.0 / math.sqrt(2.0 * math.pi)


def gelu_parts(p):
    """phi(p), phi'(p), phi''(p) of the exact-erf GELU."""
    cdf = 0.5 * (1.0 + torch.erf(p * _SQRT1_2))
    pdf = torch.exp(-0.5 * p * p) * _INV_SQRT_2PI
    return p * cdf, cdf + p * pdf, pdf * (2.0 - p * p)


def step_first_order(theta, Yb, Xn, scale, mask, dt, d, want_inputs=False):
    """Forward + first-order backward of one inner step.  Returns g_theta and the saved activations."""
    W1, b1, W2, b2, gam, bet = split_theta(theta, dt, d)
    B = Yb.shape[0]
    sv = {}
    p = Yb @ W1.t() + b1
    h, dphi, ddphi = gelu_parts(p)
    f = h @ W2.t() + b2
    r = (f * mask if mask is not None else f) + p
    mu = r.mean(1, keepdim=True)
    c = r - mu
    rstd = torch.rsqrt((c * c).mean(1, keepdim=True) + LN_EPS)
    rhat = c * rstd
    z = gam * rhat + bet
    nz = z.norm(dim=1, keepdim=True)
    yn = z / nz
    S = scale * (Xn @ yn.t())
    lse_r = torch.logsumexp(S, dim=1, keepdim=True)
    lse_c = torch.logsumexp(S, dim=0, keepdim=True)
    diag = torch.diagonal(S)
    loss = ((lse_r.squeeze(1) - diag).sum() + (lse_c.squeeze(0) - diag).sum()) / (2 * B)
    Pr = torch.exp(S - lse_r)
    Pc = torch.exp(S - lse_c)
    G = (Pr + Pc - 2 * torch.eye(B, dtype=S.dtype)) / (2 * B)
    dyn = scale * (G.t() @ Xn)
    q = (yn * dyn).sum(1, keepdim=True)
    dz = (dyn - yn * q) / nz
    dgam = (dz * rhat).sum(0)
    dbet = dz.sum(0)
    drhat = dz * gam
    m1 = drhat.mean(1, keepdim=True)
    m2 = (drhat * rhat).mean(1, keepdim=True)
    dr = rstd * (drhat - m1 - rhat * m2)
    df = dr * mask if mask is not None else dr
    dW2 = df.t() @ h
    db2 = df.sum(0)
    dh = df @ W2
    dp = dh * dphi + dr
    dW1 = dp.t() @ Yb
    db1 = dp.sum(0)
    g = torch.cat([dW1.reshape(-1), db1, dW2.reshape(-1), db2, dgam, dbet])
    sv.update(p=p, h=h, dphi=dphi, ddphi=ddphi, rstd=rstd, rhat=rhat, nz=nz, yn=yn, S=S, Pr=Pr, Pc=Pc, G=G,
              dyn=dyn, q=q, dz=dz, drhat=drhat, m1=m1, m2=m2, dr=dr, df=df, dh=dh, dp=dp, loss=loss)
    if want_inputs:
        sv["dY"] = dp @ W1
        sv["dXn"] = scale * (G @ yn)
        sv["dscale"] = (G * S).sum() / scale
    return g, sv


def step_tangent(theta, v, Yb, Xn, scale, mask, dt, d, sv):
    """Directional derivative along theta_dot = v of the WHOLE first-order step (forward-over-reverse).

    Returns (Hv [P], dY_dot [B,dt], dXn_dot [B,d], dscale_dot, L_dot) where L_dot = <g, v>.
    """
    W1, b1, W2, b2, gam, bet = split_theta(theta, dt, d)
    V1, c1, V2, c2, gamd, betd = split_theta(v, dt, d)
    B = Yb.shape[0]
    p, h, dphi, ddphi = sv["p"], sv["h"], sv["dphi"], sv["ddphi"]
    rstd, rhat, nz, yn = sv["rstd"], sv["rhat"], sv["nz"], sv["yn"]
    # tangent forward
    pd = Yb @ V1.t() + c1
    hd = dphi * pd
    fd = hd @ W2.t() + h @ V2.t() + c2
    rd = (fd * mask if mask is not None else fd) + pd
    t = (rhat * rd).mean(1, keepdim=True)
    rhatd = rstd * (rd - rd.mean(1, keepdim=True) - rhat * t)
    zd = gamd * rhat + gam * rhatd + betd
    nzd = (yn * zd).sum(1, keepdim=True)
    ynd = (zd - yn * nzd) / nz
    Sd = scale * (Xn @ ynd.t())
    G, Pr, Pc, S = sv["G"], sv["Pr"], sv["Pc"], sv["S"]
    Ld = (G * Sd).sum()
    rho = (Pr * Sd).sum(1, keepdim=True)
    kap = (Pc * Sd).sum(0, keepdim=True)
    Gd = (Pr * (Sd - rho) + Pc * (Sd - kap)) / (2 * B)
    # tangent backward
    dynd = scale * (Gd.t() @ Xn)
    dXnd = scale * (Gd @ yn + G @ ynd)
    dscaled = ((Gd * S).sum() + Ld) / scale
    dyn, q, dz = sv["dyn"], sv["q"], sv["dz"]
    qd = (ynd * dyn).sum(1, keepdim=True) + (yn * dynd).sum(1, keepdim=True)
    dzd = (dynd - ynd * q - yn * qd) / nz - dz * (nzd / nz)
    dgamd = (dzd * rhat + dz * rhatd).sum(0)
    dbetd = dzd.sum(0)
    drhat, m1, m2, dr = sv["drhat"], sv["m1"], sv["m2"], sv["dr"]
    drhatd = gamd * dz + gam * dzd
    m1d = drhatd.mean(1, keepdim=True)
    m2d = (drhatd * rhat + drhat * rhatd).mean(1, keepdim=True)
    drd = -(rstd * t) * dr + rstd * (drhatd - m1d - rhatd * m2 - rhat * m2d)
    df, dh, dp = sv["df"], sv["dh"], sv["dp"]
    dfd = drd * mask if mask is not None else drd
    dW2d = dfd.t() @ h + df.t() @ hd
    db2d = dfd.sum(0)
    dhd = dfd @ W2 + df @ V2
    dpd = dhd * dphi + dh * ddphi * pd + drd
    dW1d = dpd.t() @ Yb
    db1d = dpd.sum(0)
    dYd = dpd @ W1 + dp @ V1
    Hv = torch.cat([dW1d.reshape(-1), db1d, dW2d.reshape(-1), db2d, dgamd, dbetd])
    return Hv, dYd, dXnd, dscaled, Ld


def unrolled_match_manual(theta0, theta_tgt, Y, U, lr, scale, perms, masks=None, dt=768, d=2304):
    """Same result as ``unrolled_match_autograd`` via the reverse sweep of SURVEY.md section 8a-D7."""
    K = perms.shape[0]
    un = U.norm(dim=1, keepdim=True)
    Xn_all = U / un
    thetas = [theta0]
    ces = []
    for k in range(K):
        idx = perms[k]
        g, sv = step_first_order(thetas[-1], Y[idx], Xn_all[idx], scale, None if masks is None else masks[k], dt, d)
        ces.append(sv["loss"])
        thetas.append(thetas[-1] - lr * g)
    diffK = thetas[-1] - theta_tgt
    num = (diffK * diffK).sum()
    d0 = theta0 - theta_tgt
    den = (d0 * d0).sum()
    loss = num / den
    a = 2.0 * diffK / den
    dY = torch.zeros_like(Y)
    dXn = torch.zeros_like(U)
    dlr = torch.zeros((), dtype=Y.dtype)
    dscale = torch.zeros((), dtype=Y.dtype)
    for k in range(K - 1, -1, -1):
        idx = perms[k]
        mask = None if masks is None else masks[k]
        g, sv = step_first_order(thetas[k], Y[idx], Xn_all[idx], scale, mask, dt, d)
        Hv, dYd, dXnd, dsd, Ld = step_tangent(thetas[k], a, Y[idx], Xn_all[idx], scale, mask, dt, d, sv)
        dlr = dlr - Ld                      # d theta_{k+1} / d lr = -g_k ;  <a, g_k> = L_dot
        dY.index_add_(0, idx, -lr * dYd)
        dXn.index_add_(0, idx, -lr * dXnd)
        dscale = dscale - lr * dsd
        a = a - lr * Hv
    dU = (dXn - Xn_all * (Xn_all * dXn).sum(1, keepdim=True)) / un
    return UnrollResult(loss, num, den, dY, dU, dlr, dscale, ces, thetas[-1])


# ----------------------------------------------------------------------------------------------
# Streaming pieces, stated separately because the CUDA library exports them separately.
# ----------------------------------------------------------------------------------------------
def flat_sgd_step_ref(theta, grad, lr):
    """distill.py:582-583: theta - syn_lr * grad."""
    return theta - lr * grad


def match_loss_ref(theta_K, theta_tgt, theta_0):
    """distill.py:588-598: (sum (theta_K - theta*)^2, sum (theta_0 - theta*)^2)."""
    return (F.mse_loss(theta_K, theta_tgt, reduction="sum"), F.mse_loss(theta_0, theta_tgt, reduction="sum"))


def momentum_sgd_ref(param, grad, buf, lr, momentum, first):
    """torch.optim.SGD(momentum=m, dampening=0, nesterov=False, weight_decay=0) as built at distill.py:233-241."""
    buf = grad.clone() if first else momentum * buf + grad
    return param - lr * buf, buf


def make_problem(N=100, B=100, K=8, dt=768, d=2304, seed=0, dtype=torch.float32, dropout=False,
                 lr=0.1, scale=2.6593, tgt_eps=0.01):
    """seeded synthetic segment of Flickr shape."""
    gen = torch.Generator().manual_seed(seed)
    lin1 = torch.nn.Linear(dt, d)
    lin2 = torch.nn.Linear(d, d)
    with torch.no_grad():
        # nn.Linear default init (kaiming_uniform(a=sqrt 5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)), seeded here
        for lin in (lin1, lin2):
            bound = 1.0 / math.sqrt(lin.in_features)
            lin.weight.copy_((torch.rand(lin.weight.shape, generator=gen) * 2 - 1) * bound)
            lin.bias.copy_((torch.rand(lin.bias.shape, generator=gen) * 2 - 1) * bound)
    theta0 = torch.cat([lin1.weight.detach().reshape(-1), lin1.bias.detach(), lin2.weight.detach().reshape(-1),
                        lin2.bias.detach(), torch.ones(d), torch.zeros(d)]).to(dtype)
    theta_tgt = theta0 + tgt_eps * torch.randn(theta0.shape, generator=gen).to(dtype)
    Y = (torch.randn(N, dt, generator=gen) * 0.5253 - 0.0094).to(dtype)     # distill_original.py:147 text-noise stats
    U = torch.randn(N, d, generator=gen).to(dtype)
    perms = torch.stack([torch.randperm(N, generator=gen)[:B] for _ in range(K)])
    masks = None
    if dropout:
        masks = ((torch.rand(K, B, d, generator=gen) >= DROPOUT_P).to(dtype) / (1.0 - DROPOUT_P))
    return dict(theta0=theta0, theta_tgt=theta_tgt, Y=Y, U=U, lr=torch.tensor(lr, dtype=dtype),
                scale=torch.tensor(scale, dtype=dtype), perms=perms, masks=masks, dt=dt, d=d)
