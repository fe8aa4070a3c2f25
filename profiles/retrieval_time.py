"""Developer aid: vldd_sim_rank_fused wall time per shape (CUDA events, best of 5).  python profiles/retrieval_time.py 1000 5000 25000"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
if os.environ.get("QUICK_LIB"):            # same-box A/B of two builds: load this library instead of the in-tree one
    import multimodal_dataset_distillation_b200.build as _b
    _b.build_library = lambda force=False, verbose=False: os.environ["QUICK_LIB"]
import bench
from multimodal_dataset_distillation_b200 import ops
dev = torch.device("cuda")
for I in [int(a) for a in sys.argv[1:]] or [5000]:
    D = 2304 if I <= 1000 else 768
    img, txt, t2i, gptr, gidx = bench.gpu_retrieval_set(I, 5, D, dev, seed=I)
    for _ in range(2):
        r1, r2 = ops.sim_rank_fused(img, txt, t2i, gptr, gidx, 14.285714)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r1, r2 = ops.sim_rank_fused(img, txt, t2i, gptr, gidx, 14.285714); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    t = min(ts)
    print(f"{I} x {5 * I} x {D}: {t:.3f} ms  {I * 5 * I / t / 1e6:.1f} G pairs/s  checksum {int(r1.sum())} {int(r2.sum())}")
