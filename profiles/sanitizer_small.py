"""Small run of every round-2 kernel for compute-sanitizer (memcheck): screened retrieval incl. overflow fall-back, the
caption-shard form, the unroll engine with in-engine dropout, fused outer update, one-pass matching loss."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import retrieval_ref as RR
from multimodal_dataset_distillation_b200 import ops, distill, dist as Dm

dev = lambda a: torch.from_numpy(a).cuda()
I, C, D = 300, 5, 64
img, txt = RR.synthetic_retrieval(I, C, D, seed=11)
img, txt = (np.round(img * 8) / 8).astype(np.float32), (np.round(txt * 8) / 8).astype(np.float32)
T = I * C
t2i, ptr, idx = ops.maps_to_arrays(*RR.flickr_maps(I, C), I, T)
f1, f2 = ops.sim_rank_fused(dev(img), dev(txt), dev(t2i), dev(ptr), dev(idx), 14.285714)
m1, m2 = ops.sim_rank(dev(img), dev(txt), dev(t2i), dev(ptr), dev(idx), 14.285714)
assert torch.equal(f1, m1) and torch.equal(f2, m2)
parts = []
for r in range(2):
    lo, hi = Dm.shard_bounds(T, 2, r)
    parts.append(ops.FusedRankShard(dev(img), dev(txt)[lo:hi].contiguous(), lo, dev(t2i)[lo:hi].contiguous(), dev(ptr), dev(idx), 14.285714))
cands = [p.candidates() for p in parts]
thr_s, thr_i = Dm.merge_candidates(torch.stack([c[0] for c in cands]), torch.stack([c[1] for c in cands]))
outs = [p.count(thr_s, thr_i, 0) for p in parts]
assert torch.equal(sum(o[0] for o in outs), f1)

N, B, K, dt, d = 48, 32, 2, 64, 96
args = distill.parse_args(["--syn_steps", str(K), "--expert_epochs", "1", "--max_start_epoch", "2", "--num_queries", str(N),
                           "--mini_batch_size", str(B), "--logit_scale_mode", "fork", "--student_dropout", "0.1"])
g = torch.Generator().manual_seed(4)
eng = distill.DistillEngine(torch.randn(N, d, generator=g), torch.randn(N, dt, generator=g),
                            distill.synthetic_experts(2, 3, dt, d, seed=1).cuda(), args, "cuda")
for i in range(3):
    loss = eng.step_fast(i % 2, i % 2)
torch.cuda.synchronize()
print("sanitizer workload ok", float(loss), int(f1.sum()))
