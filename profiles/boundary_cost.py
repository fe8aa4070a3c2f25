"""Developer aid: what the per-iteration launches outside the engine's graph cost.  ms per iteration of (a) the engine call alone,
(b) engine call + fused outer update (= DistillEngine.step_fast at N = 1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodal_dataset_distillation_b200 import distill, ops

args = bench.bench_args()
U, Y = bench.make_pairs(0)
eng = distill.DistillEngine(U, Y, bench.make_experts(1).cuda(), args, "cuda")
K, B, N = bench.CFG["K"], bench.CFG["B"], bench.CFG["N"]
g = torch.Generator().manual_seed(0)
perms = [torch.stack([torch.randperm(N, generator=g)[:B] for _ in range(K)]).cuda() for _ in range(8)]

def engine_only(i):
    e, s = i % 4, (i // 4) % 2
    ops.unrolled_match(eng.experts[e, s], eng.experts[e, s + 1], eng.Y.detach(), eng.U.detach(), eng.syn_lr_txt.detach(),
                       eng.syn_lr_img.detach(), perms[i % 8], None, eng.ws, dropout_p=0.1, rng_state=eng.rng_state, clone_results=False)

def full(i):
    eng.step_fast(i % 4, (i // 4) % 2, perms[i % 8])

for name, fn in (("engine call only", engine_only), ("engine call + outer update (step_fast)", full), ("engine call only", engine_only)):
    for i in range(20):
        fn(i)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(200):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 200)
    print(f"{name:42s} {best:.4f} ms/iter")
