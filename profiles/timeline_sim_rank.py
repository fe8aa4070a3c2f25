"""Developer aid: kernel timeline (CUPTI via torch.profiler) of one vldd_sim_rank call at the Flickr test shape."""
import os, sys, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from multimodal_dataset_distillation_b200 import ops
from oracle import retrieval_ref as RR
I, C, D = 1000, 5, 768
T = I * C
img, txt = RR.synthetic_retrieval(I, C, D, seed=0)
txt2img, img2txt = RR.flickr_maps(I, C)
t2i, ptr, idx = ops.maps_to_arrays(txt2img, img2txt, I, T)
d = lambda a: torch.from_numpy(a).cuda()
img_d, txt_d, t2i_d, ptr_d, idx_d = d(img), d(txt), d(t2i), d(ptr), d(idx)
ws = torch.empty(ops.lib().vldd_sim_rank_workspace_bytes(I, T, D), dtype=torch.uint8, device="cuda")
for _ in range(5):
    ops.sim_rank(img_d, txt_d, t2i_d, ptr_d, idx_d, 14.285714, ws)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        ops.sim_rank(img_d, txt_d, t2i_d, ptr_d, idx_d, 14.285714, ws)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = sorted([e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")], key=lambda e: e["ts"])
t0 = ev[0]["ts"]
for e in ev:
    print(f"{e['ts'] - t0:9.1f} {e['dur']:7.2f}  {e['name'][:110]}")
