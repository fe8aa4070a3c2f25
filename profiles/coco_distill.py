"""BASELINE.json configs[3]: COCO-shaped distillation (N = 500 pairs; minibatch 100 = true random subsets, and 500) on one GPU.
    python profiles/coco_distill.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodal_dataset_distillation_b200 import distill

def run(N, B, K=8, steps=20):
    args = distill.parse_args(["--syn_steps", str(K), "--expert_epochs", "1", "--max_start_epoch", "2", "--num_queries", str(N),
                               "--mini_batch_size", str(B), "--lr_img", "1000", "--lr_txt", "1000", "--lr_lr", "0.01",
                               "--logit_scale_mode", "upstream", "--student_dropout", "0.1"])
    g = torch.Generator().manual_seed(0)
    U = torch.randn(N, 2304, generator=g)
    Y = torch.randn(N, 768, generator=g) * 0.5253 - 0.0094
    eng = distill.DistillEngine(U, Y, bench.make_experts(1).cuda(), args, "cuda")
    perms = [torch.stack([torch.randperm(N, generator=g)[:B] for _ in range(K)]).cuda() for _ in range(4)]
    for i in range(3):
        eng.outer_step(eng.segment_loss(i % 4, 0, perms[i % 4]))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = eng.segment_loss(i % 4, (i // 4) % 2, perms[i % 4])
        eng.outer_step(loss)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"N={N} B={B} K={K}: {ms:.3f} ms / iteration = {1e3 / ms:.1f} it/s   loss={float(loss):.6f}", flush=True)

if __name__ == "__main__":
    run(100, 100)
    run(500, 100)
    run(500, 500)
