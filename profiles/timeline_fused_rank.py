"""Developer aid: kernel timeline (CUPTI via torch.profiler) of one vldd_sim_rank_fused call.  python profiles/timeline_fused_rank.py 5000"""
import os, sys, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from multimodal_dataset_distillation_b200 import ops
I = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
img, txt, t2i, gptr, gidx = bench.gpu_retrieval_set(I, 5, 768, torch.device("cuda"), seed=I)
for _ in range(3):
    ops.sim_rank_fused(img, txt, t2i, gptr, gidx, 14.285714)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ops.sim_rank_fused(img, txt, t2i, gptr, gidx, 14.285714)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = sorted([e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")], key=lambda e: e["ts"])
t0 = ev[0]["ts"]
for e in ev:
    print(f"{e['ts'] - t0:9.1f} {e['dur']:8.2f}  {e['name'].replace('vldd::', '')[:100]}")
print(f"# span {ev[-1]['ts'] + ev[-1]['dur'] - t0:.1f} us")
