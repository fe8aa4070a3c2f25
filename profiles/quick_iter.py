"""Developer aid: ms per distill iteration (device-resident segments, rotating over 4 experts x 2 start epochs), nothing else.
python profiles/quick_iter.py [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
if os.environ.get("QUICK_LIB"):            # same-box A/B of two builds: load this library instead of the in-tree one
    import multimodal_dataset_distillation_b200.build as _b
    _b.build_library = lambda force=False, verbose=False: os.environ["QUICK_LIB"]
import bench
from multimodal_dataset_distillation_b200 import distill

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
args = bench.bench_args()
U, Y = bench.make_pairs(0)
eng = distill.DistillEngine(U, Y, bench.make_experts(1).cuda(), args, "cuda")
for i in range(20):
    eng.step_fast(i % 4, i % 2)
torch.cuda.synchronize()
best = []
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        eng.step_fast(i % 4, (i // 4) % 2)
    b.record()
    torch.cuda.synchronize()
    best.append(a.elapsed_time(b) / iters)
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("VLDD_"))
print(f"{min(best):.4f} ms/iter (runs: {', '.join(f'{x:.4f}' for x in best)})  loss={float(eng.ws.out5[2]):.6f}  [{tag}]")
