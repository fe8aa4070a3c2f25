"""Generate tests/golden/*.json|npz by running the REFERENCE's own importable code on seeded inputs.

Run once in the build container (needs /root/reference; the GPU box does not have it):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

What is real reference code here:
  * /root/reference/epoch.py:219           itm_eval (fork)
  * /root/reference/epoch_original.py:115  itm_eval (upstream; ``utils`` stubbed -- it only provides MetricLogger)
  * /root/reference/epoch_original.py:68   epoch_test, driven with a fake model whose image encoder is the identity
  * /root/reference/reparam_module.py      ReparamModule (flat-param layout + functional forward)
networks.ProjectionHead (networks.py:625-646) is unimportable (module import downloads BERT), so the same
nn.Module is declared below with identical attribute names / order; the inner loop of distill.py:509-606 is
driven through the real ReparamModule + torch double-backward.
"""
import contextlib
import io
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
REF = "/root/reference"
sys.path.insert(0, REF)

from oracle import retrieval_ref as RR  # noqa: E402  (input generators only)
from oracle import distill_ref as DR    # noqa: E402  (problem generator only)


def _import_reference():
    import epoch as ref_epoch                                    # fork
    stub = types.ModuleType("utils")

    class MetricLogger:                                          # epoch_original.py:71 only constructs it
        def __init__(self, *a, **k):
            pass
    stub.MetricLogger = MetricLogger
    saved = sys.modules.get("utils")
    sys.modules["utils"] = stub
    import epoch_original as ref_epoch_orig
    if saved is not None:
        sys.modules["utils"] = saved
    else:
        del sys.modules["utils"]
    import reparam_module as ref_reparam
    return ref_epoch, ref_epoch_orig, ref_reparam


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


class ProjectionHead(nn.Module):
    """Same module tree as networks.py:625-646 (attribute names and registration order matter for the flat layout)."""

    def __init__(self, embedding_dim, projection_dim=768, dropout=0.1):
        super().__init__()
        self.projection = nn.Linear(embedding_dim, projection_dim)
        self.gelu = nn.GELU()
        self.fc = nn.Linear(projection_dim, projection_dim)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(projection_dim)

    def forward(self, x):
        projected = self.projection(x)
        x = self.gelu(projected)
        x = self.fc(x)
        x = self.dropout(x)
        x = x + projected
        return self.layer_norm(x)


def golden_retrieval(ref_epoch, ref_epoch_orig):
    cases = []
    specs = [  # (name, I, C, D, seed, quantise, topk_fill)
        ("tiny", 7, 5, 16, 0, None, False),
        ("small", 40, 5, 32, 1, None, False),
        ("c1_roco", 64, 1, 24, 2, None, False),
        ("ties_q64", 48, 5, 32, 3, 64.0, False),
        ("fill128", 60, 5, 48, 4, None, True),
        ("flickr_768", 1000, 5, 768, 0, None, False),
        ("flickr_768_fill", 1000, 5, 768, 0, None, True),
    ]
    for name, I, C, D, seed, quant, fill in specs:
        img, txt = RR.synthetic_retrieval(I, C, D, seed=seed)
        S = (np.float32(14.285714) * img) @ txt.T
        if quant:
            S = np.round(S * quant) / np.float32(quant)
        S = S.astype(np.float32)
        St = np.ascontiguousarray(S.T)
        if fill:
            S, St = RR.topk_fill_ref(S, 128), RR.topk_fill_ref(St, 128)
        txt2img, img2txt = RR.flickr_maps(I, C)
        fork = quiet(ref_epoch.itm_eval, S, St, txt2img, img2txt)
        orig = quiet(ref_epoch_orig.itm_eval, S, St, txt2img, img2txt)
        case = dict(name=name, I=I, C=C, D=D, seed=seed, quant=quant, fill=fill,
                    fork={k: float(v) for k, v in fork.items()}, orig={k: float(v) for k, v in orig.items()},
                    score_checksum=float(np.float64(S).sum()))
        if not quant and not fill:
            # tie-free: the reference's ranks are well defined -> record them (recomputed with the reference's own lines)
            r_i = [int(np.min(np.where(np.isin(np.argsort(row)[::-1], img2txt[i]))[0])) for i, row in enumerate(S)]
            r_t = [int(np.where(np.argsort(row)[::-1] == txt2img[t])[0][0]) for t, row in enumerate(St)]
            case["ranks_i2t"], case["ranks_t2i"] = r_i, r_t
        cases.append(case)
        print("retrieval", name, fork["r_mean"], orig["r_mean"])
    return cases


def golden_epoch_test(ref_epoch_orig):
    """Drive the reference's epoch_test (epoch_original.py:68-111) with a fake model on CPU."""
    I, C, D, dt = 150, 5, 24, 12      # T=750 > 128 so the top-128/-100 fill is exercised in both directions
    gen = torch.Generator().manual_seed(5)
    head = ProjectionHead(dt, D)
    with torch.no_grad():
        for p in head.parameters():
            p.copy_(torch.randn(p.shape, generator=gen) * 0.2)
    feats = torch.randn(I, D, generator=gen)
    bert = torch.randn(I * C, dt, generator=gen)

    class FakeModel(nn.Module):
        def __init__(self):
            super().__init__()
            self.text_projection = head
            self.image_encoder = nn.Identity()

    loader = [(feats[i:i + 16], torch.arange(i, min(i + 16, I))) for i in range(0, I, 16)]
    orig_to = torch.Tensor.to

    def to_cpu(self, *a, **k):        # epoch_original.py:77,95,102 hard-code .to('cuda'); run it on CPU
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        return orig_to(self, *a, **k)
    torch.Tensor.to = to_cpu
    try:
        s_i2t, s_t2i = quiet(ref_epoch_orig.epoch_test, loader, FakeModel(), "cpu", bert)
    finally:
        torch.Tensor.to = orig_to
    theta = torch.cat([p.detach().reshape(-1) for p in head.parameters()])
    np.savez_compressed(os.path.join(HERE, "epoch_test_small.npz"), theta=theta.numpy(), feats=feats.numpy(),
                        bert=bert.numpy(), s_i2t=s_i2t, s_t2i=s_t2i, dims=np.array([I, C, D, dt]))
    print("epoch_test", s_i2t.shape, s_t2i.shape, float((s_i2t > -100).sum()))


def golden_reparam(ref_reparam):
    torch.manual_seed(0)
    head = ProjectionHead(768, 2304)
    rp = ref_reparam.ReparamModule(head)
    info = dict(param_numel=int(rp.param_numel), names=[f"{mn}.{n}" for mn, n in rp._param_infos],
                numels=[int(x) for x in rp._param_numels], shapes=[list(s) for s in rp._param_shapes])
    # functional forward through the real class on a small head, eval mode (dropout off)
    torch.manual_seed(1)
    small = ProjectionHead(12, 20)
    rps = ref_reparam.ReparamModule(small).eval()
    theta = torch.randn(rps.param_numel) * 0.3
    x = torch.randn(5, 12)
    out = rps(x, flat_param=theta)
    out2 = rps(x, flat_param=theta.unsqueeze(0))       # DataParallel [1,P] convention, reparam_module.py:149
    assert torch.equal(out, out2)
    np.savez_compressed(os.path.join(HERE, "reparam_small.npz"), theta=theta.numpy(), x=x.numpy(), out=out.detach().numpy())
    print("reparam", info["param_numel"], info["names"])
    return info


def run_reference_unroll(ref_reparam, pr, dt, d, train_mode_masks):
    """distill.py:509-606 through the real ReparamModule (text tower; image side = embeddings U)."""
    head = ProjectionHead(dt, d)
    net = ref_reparam.ReparamModule(head)
    net.train()                                                    # distill.py:447
    if pr["masks"] is None:
        head.dropout.p = 0.0                                       # parity mode: dropout disabled
    Y = pr["Y"].clone().requires_grad_(True)
    U = pr["U"].clone().requires_grad_(True)
    lr = pr["lr"].clone().requires_grad_(True)
    scale = pr["scale"].clone().requires_grad_(True)
    params = [pr["theta0"].clone().requires_grad_(True)]
    ces = []
    for k in range(pr["perms"].shape[0]):
        idx = pr["perms"][k]
        x = U[idx]
        x = x / x.norm(dim=1, keepdim=True)
        if pr["masks"] is not None:
            # inject the mask: replace dropout by multiplication with the recorded (pre-scaled) mask
            mk = pr["masks"][k]
            head.dropout = _MaskMul(mk)
        y = net(Y[idx], flat_param=params[-1])
        y = y / y.norm(dim=1, keepdim=True)
        logits = scale * x.float() @ y.float().t() if x.dtype == torch.float32 else scale * x @ y.t()
        gt = torch.arange(len(logits))
        ce = (F.cross_entropy(logits, gt) + F.cross_entropy(logits.t(), gt)) / 2
        ces.append(float(ce))
        g = torch.autograd.grad(ce, params[-1], create_graph=True)[0]
        params.append(params[-1] - lr * g)
    num = F.mse_loss(params[-1], pr["theta_tgt"], reduction="sum")
    den = F.mse_loss(pr["theta0"], pr["theta_tgt"], reduction="sum")
    loss = num / den
    loss.backward()
    return dict(loss=float(loss), num=float(num), den=float(den), dlr=float(lr.grad), dscale=float(scale.grad),
                ce=ces, dY=Y.grad.detach(), dU=U.grad.detach(), theta_K=params[-1].detach())


class _MaskMul(nn.Module):
    def __init__(self, mask):
        super().__init__()
        self.mask = mask

    def forward(self, x):
        return x * self.mask


def golden_distill(ref_reparam):
    out = {}
    # small cases: every tensor stored (fp64 = truth, fp32 = reference behaviour)
    small = {}
    for name, kw in {
        "small_nodrop": dict(N=12, B=8, K=3, dt=10, d=16, seed=1, lr=0.3, scale=2.0, tgt_eps=0.05),
        "small_drop": dict(N=12, B=8, K=3, dt=10, d=16, seed=2, lr=0.3, scale=2.0, tgt_eps=0.05, dropout=True),
        "mid_full_batch": dict(N=24, B=24, K=4, dt=32, d=48, seed=3, lr=0.2, scale=14.2857, tgt_eps=0.02),
    }.items():
        for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            pr = DR.make_problem(dtype=dtype, **kw)
            res = run_reference_unroll(ref_reparam, pr, kw["dt"], kw["d"], None)
            small[f"{name}_{tag}_dY"] = res["dY"].numpy()
            small[f"{name}_{tag}_dU"] = res["dU"].numpy()
            small[f"{name}_{tag}_thetaK"] = res["theta_K"].numpy()
            out[f"{name}_{tag}"] = dict(kw=kw, loss=res["loss"], num=res["num"], den=res["den"], dlr=res["dlr"],
                                        dscale=res["dscale"], ce=res["ce"])
            print("distill", name, tag, res["loss"], res["dlr"], res["dscale"])
    np.savez_compressed(os.path.join(HERE, "distill_small.npz"), **small)
    # Flickr-shape config 3 (N=B=100, K=8, 768->2304): scalars + sampled entries + norms, fp32 and fp64
    flick = {}
    for scale_name, scale in (("upstream", 2.6593), ("fork", 0.1), ("eval", 14.2857)):
        for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            if tag == "f32" and scale_name != "upstream":
                continue
            kw = dict(N=100, B=100, K=8, dt=768, d=2304, seed=0, lr=0.1, scale=scale, tgt_eps=0.01)
            pr = DR.make_problem(dtype=dtype, **kw)
            res = run_reference_unroll(ref_reparam, pr, 768, 2304, None)
            key = f"flickr_{scale_name}_{tag}"
            # stored rounded to fp32 (quantisation 6e-8 relative, far below the 1e-4 tolerance); only the
            # upstream/f64 case keeps every row, the others keep every 10th row to stay small
            rows = slice(None) if key == "flickr_upstream_f64" else slice(None, None, 10)
            flick[key + "_dY"] = res["dY"].numpy()[rows].astype(np.float32)
            flick[key + "_dU"] = res["dU"].numpy()[rows].astype(np.float32)
            tk = res["theta_K"].double()
            out[key] = dict(kw=kw, loss=res["loss"], num=res["num"], den=res["den"], dlr=res["dlr"], dscale=res["dscale"],
                            ce=res["ce"], thetaK_sum=float(tk.sum()), thetaK_sqsum=float((tk * tk).sum()),
                            thetaK_sample=[float(x) for x in tk[:: 700001]])
            print("distill", key, res["loss"], res["dlr"], res["dscale"])
    # one dropout case at Flickr shape (fp64 only) with injected masks
    kw = dict(N=100, B=100, K=2, dt=768, d=2304, seed=7, lr=0.1, scale=2.6593, tgt_eps=0.01, dropout=True)
    pr = DR.make_problem(dtype=torch.float64, **kw)
    res = run_reference_unroll(ref_reparam, pr, 768, 2304, None)
    flick["flickr_drop_f64_dY"] = res["dY"].numpy()[::10].astype(np.float32)
    flick["flickr_drop_f64_dU"] = res["dU"].numpy()[::10].astype(np.float32)
    out["flickr_drop_f64"] = dict(kw=kw, loss=res["loss"], num=res["num"], den=res["den"], dlr=res["dlr"],
                                  dscale=res["dscale"], ce=res["ce"])
    np.savez_compressed(os.path.join(HERE, "distill_flickr.npz"), **flick)
    return out


def main():
    torch.set_num_threads(8)
    ref_epoch, ref_epoch_orig, ref_reparam = _import_reference()
    gold = dict(
        generated_by="tests/golden/make_golden.py", torch=torch.__version__, numpy=np.__version__,
        retrieval=golden_retrieval(ref_epoch, ref_epoch_orig),
        reparam=golden_reparam(ref_reparam),
        distill=golden_distill(ref_reparam),
    )
    golden_epoch_test(ref_epoch_orig)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", os.path.join(HERE, "golden.json"))


if __name__ == "__main__":
    main()
