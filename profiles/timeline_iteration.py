"""Developer aid: REAL (concurrent, graph-replayed) kernel timeline of one distill iteration from CUPTI activity records
(torch.profiler; no serialisation, unlike ncu).  python profiles/timeline_iteration.py > gpurun_out/timeline.txt"""
import os, sys, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from multimodal_dataset_distillation_b200 import distill

args = bench.bench_args()
U, Y = bench.make_pairs(0)
eng = distill.DistillEngine(U, Y, bench.make_experts(1).cuda(), args, "cuda")
for i in range(12):
    eng.step_fast(i % 4, i % 2)
torch.cuda.synchronize()
SYNC = os.environ.get("TIMELINE_SYNC", "0") == "1"        # 1: synchronise after every iteration (exposes the host's launch time)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(5):
        eng.step_fast(i % 4, i % 2)
        if SYNC:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
# one iteration = from a set_stage_sources_kernel (first kernel of vldd_unrolled_match, outside the graph) to the next one
starts = [i for i, e in enumerate(ev) if "set_stage_sources_kernel" in e["name"]]
it = ev[starts[-2]:starts[-1]] if not SYNC else ev[starts[-1]:]
t0 = it[0]["ts"]
end = max(e["ts"] + e["dur"] for e in it)
print(f"# {len(it)} activities, span {end - t0:.1f} us")
busy, last = 0.0, t0
gaps = []
for e in it:
    s, f = e["ts"], e["ts"] + e["dur"]
    if s > last:
        gaps.append((s - last, e["name"][:60], s - t0))
    if f > last:
        busy += f - max(s, last); last = f
print(f"# union busy {busy:.1f} us, idle {end - t0 - busy:.1f} us in {len(gaps)} gaps")
def short(n):
    n = n.replace("vldd::", "").replace("void ", "")
    return n[:78]
print("# start_us  dur_us stream  kernel")
for e in it:
    print(f"{e['ts'] - t0:9.1f} {e['dur']:7.2f} {e['args'].get('stream', -1):5}  {short(e['name'])}")
agg = {}
for e in it:
    k = short(e["name"]); a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e["dur"]
print("# per kernel (concurrent durations)")
for k, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{d:9.1f} us x{c:3d}  avg {d / c:6.2f}  {k}")
