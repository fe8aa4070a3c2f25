"""Developer aid: per-call-site GPU time of one distill iteration (serialised, warm L2): VLDD_PROFILE=1 python profiles/profile_iteration.py"""
import os, sys
os.environ["VLDD_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodal_dataset_distillation_b200 import distill
args = bench.bench_args()
U, Y = bench.make_pairs(0)
eng = distill.DistillEngine(U, Y, bench.make_experts(1).cuda(), args, "cuda")
for i in range(4):
    sys.stderr.write(f"---- iteration {i}\n")
    loss = eng.segment_loss(i % 4, 0)
    torch.cuda.synchronize()
