"""Small end-to-end case touching every family of kernels once, without CUDA graphs (plain launches, easy to put under a
debugger or a profiler):    python profiles/small_end_to_end.py
(compute-sanitizer is closed on this GPU pool; the checks here are comparisons with the oracle and between code paths.)
Covers the tcgen05 GEMM paths (dims satisfy the tensor-map constraints), the row kernels, InfoNCE, retrieval (materialised
and fused), the clip loss, the Mode-B node and the nearest-row lookup."""
import os, sys
os.environ.setdefault("VLDD_GRAPH", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import distill_ref as R, retrieval_ref as RR
from multimodal_dataset_distillation_b200 import ops, epoch, infonce

pr = R.make_problem(N=40, B=32, K=2, dt=64, d=128, seed=0, dropout=True)
c = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in pr.items()}
res = ops.unrolled_match(c["theta0"], c["theta_tgt"], c["Y"], c["U"], c["lr"], c["scale"], c["perms"], c["masks"])
ref = R.unrolled_match_manual(**{k: (v.double() if isinstance(v, torch.Tensor) and v.is_floating_point() else v) for k, v in pr.items()})
err = float((res["dY"].cpu().double() - ref.dY).abs().max() / ref.dY.abs().max())
print("unrolled_match dY rel err", err)
got = ops.clip_loss(c["theta0"], c["Y"][:32].contiguous(), c["U"][:32].contiguous())
print("clip_loss", float(got["loss"]), got["top1"].tolist())
img, txt = RR.synthetic_retrieval(160, 5, 64, seed=1)
t2i, ptr, idx = ops.maps_to_arrays(*RR.flickr_maps(160, 5), 160, 800)
d = lambda a: torch.from_numpy(a).cuda()
r1, r2 = ops.sim_rank(d(img), d(txt), d(t2i), d(ptr), d(idx))
f1, f2 = ops.sim_rank_fused(d(img), d(txt), d(t2i), d(ptr), d(idx))
print("sim_rank == fused:", bool(torch.equal(r1, f1) and torch.equal(r2, f2)), epoch.ranks_to_result(r1, r2)["r_mean"])
x = torch.randn(32, 128, device="cuda", requires_grad=True)
y = torch.randn(32, 128, device="cuda", requires_grad=True)
loss = infonce.infonce_loss(x, y, 2.0)
gx, = torch.autograd.grad(loss, x, create_graph=True)
(gx.pow(2).sum()).backward()
print("mode B double backward ok", float(y.grad.abs().sum()))
print("nearest", ops.nearest_rows(d(txt[:8].copy()), d(txt)).tolist())
torch.cuda.synchronize()
assert err < 1e-4
