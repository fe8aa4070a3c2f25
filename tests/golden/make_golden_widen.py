"""Goldens for the two "next" rows of SURVEY.md section 8f, made by EXECUTING the reference's own source in the build
container (nothing is copied into this repository; the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_widen.py

  clip_forward.npz  CLIPModel_full.forward (networks.py:845-889): the method's source is cut out of networks.py with `ast`
                    (the module itself cannot be imported: it downloads BERT at import time) and bound to a stand-in object
                    whose image encoder is the identity and whose text_projection is a ProjectionHead with the reference's
                    module tree.  Outputs: loss, acc and, through loss.backward(), the gradients of the head parameters,
                    the text features and the image features (fp32: the method casts its inputs with .float()).
  nearest.npz       nearest_neighbor (distill.py:89-95), cut out the same way; it calls sklearn's cosine_similarity.
"""
import ast
import contextlib
import io
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
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True
REF = "/root/reference"

from make_golden import ProjectionHead, _MaskMul   # noqa: E402  (reference module tree, see make_golden.py)
from oracle import distill_ref as DR                # noqa: E402  (problem generator only)
from oracle import retrieval_ref as RR              # noqa: E402  (problem generator only)


def cut_function(path, name, cls=None):
    """Source text of a module-level function or of method `cls.name`, dedented, straight from the reference file."""
    src = open(path).read()
    tree = ast.parse(src)
    nodes = tree.body
    if cls is not None:
        nodes = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    fn = next(n for n in nodes if isinstance(n, ast.FunctionDef) and n.name == name)
    lines = src.splitlines()[fn.lineno - 1:fn.end_lineno]
    indent = len(lines[0]) - len(lines[0].lstrip())
    return "\n".join(l[indent:] for l in lines)


class _Identity(nn.Module):
    def forward(self, x):
        return x


def clip_forward_golden():
    ns = {"torch": torch, "np": np, "F": F}
    exec(cut_function(os.path.join(REF, "networks.py"), "forward", cls="CLIPModel_full"), ns)
    ref_forward = ns["forward"]
    out = {}
    # the reference casts the features with .float() (networks.py:868-870), so its own forward only exists in fp32; the
    # fp64 truth the GPU tests compare against is the oracle restatement, itself checked against these fp32 outputs
    for tag, B, dt, d, drop, dtype in (("small", 12, 16, 24, False, torch.float32), ("drop", 20, 24, 40, True, torch.float32),
                                       ("flickr", 100, 768, 2304, True, torch.float32)):
        pr = DR.make_problem(N=B, B=B, K=1, dt=dt, d=d, seed=21, dropout=drop)
        head = ProjectionHead(dt, d).to(dtype)
        with torch.no_grad():                                  # parameters() order == ReparamModule flat order
            off = 0
            for p_ in head.parameters():
                p_.copy_(pr["theta0"][off:off + p_.numel()].reshape(p_.shape))
                off += p_.numel()
        if drop:
            head.dropout = _MaskMul(pr["masks"][0].to(dtype))
        else:
            head.eval()
        obj = types.SimpleNamespace(image_encoder=_Identity(), text_encoder=_Identity(), text_projection=head, distill=True)
        Y = pr["Y"].detach().clone().to(dtype).requires_grad_(True)
        # image features correlated with the projected captions, so that the top-1 counters are exercised
        with torch.no_grad():
            z = head(Y)
            U0 = z + (1.5 if d <= 64 else 12.0) * z.std() * torch.randn(z.shape, generator=torch.Generator().manual_seed(77), dtype=dtype)
        U = U0.clone().requires_grad_(True)
        with contextlib.redirect_stdout(io.StringIO()):
            loss, acc = ref_forward(obj, U, Y, 0)
        loss.backward()
        g_theta = torch.cat([p_.grad.reshape(-1) for p_ in head.parameters()])
        small = d <= 64
        out[f"{tag}_U"] = U0.numpy()                           # theta, Y, mask: DR.make_problem(seed=21) in the tests
        out[f"{tag}_loss"] = np.array(float(loss.detach()))
        out[f"{tag}_acc"] = np.array(float(acc))
        out[f"{tag}_dY"] = Y.grad.numpy()
        out[f"{tag}_dU"] = U.grad.numpy() if small else U.grad.numpy()[::7].copy()
        out[f"{tag}_g_theta"] = g_theta.numpy() if small else g_theta.numpy()[::997].copy()   # strided sample at full size
    np.savez_compressed(os.path.join(HERE, "clip_forward.npz"), **out)
    print("clip_forward.npz:", {k: (v.shape if v.ndim else float(v)) for k, v in out.items() if k.endswith(("loss", "acc"))})


def nearest_golden():
    from sklearn.metrics.pairwise import cosine_similarity
    ns = {"np": np, "cosine_similarity": cosine_similarity}
    exec(cut_function(os.path.join(REF, "distill.py"), "nearest_neighbor"), ns)
    ref_nn = ns["nearest_neighbor"]
    out = {}
    for tag in ("small", "mid"):
        query, bank = RR.nearest_problem(tag)       # seeded generator shared with the tests (inputs are not stored)
        sentences = list(range(bank.shape[0]))      # "sentence" = its own index
        got = ref_nn(sentences, torch.from_numpy(query), bank)
        out[f"{tag}_idx"] = np.asarray(got, dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "nearest.npz"), **out)
    print("nearest.npz:", out["small_idx"], out["mid_idx"][:10])


if __name__ == "__main__":
    clip_forward_golden()
    nearest_golden()
