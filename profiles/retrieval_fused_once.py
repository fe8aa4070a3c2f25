"""One vldd_sim_rank_fused call per shape (developer aid for `ncu --metrics gpu__time_duration.sum`: per-pass kernel times)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodal_dataset_distillation_b200 import ops
dev = torch.device("cuda")
for I in [int(a) for a in sys.argv[1:]] or [5000]:
    img, txt, t2i, gptr, gidx = bench.gpu_retrieval_set(I, 5, 768, dev, seed=I)
    for _ in range(2):
        r1, r2 = ops.sim_rank_fused(img, txt, t2i, gptr, gidx, 14.285714)
    torch.cuda.synchronize()
    print(I, int(r1.sum()), int(r2.sum()))
