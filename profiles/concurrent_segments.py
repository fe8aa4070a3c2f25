"""Throughput mode on ONE GPU: G expert segments of one outer step in flight at once (one CUDA stream + one engine
workspace each), gradients summed before the outer update -- the same G-segment minibatch the multi-GPU run computes,
here used to fill the latency-bound gaps of a single segment's launch chain.    python profiles/concurrent_segments.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodal_dataset_distillation_b200 import distill, ops

def run(G, steps=20):
    args = bench.bench_args()
    U, Y = bench.make_pairs(0)
    experts = bench.make_experts(100).cuda()
    engs = [distill.DistillEngine(U, Y, experts, args, "cuda") for _ in range(G)]
    for e in engs[1:]:                                  # all segments differentiate the SAME synthetic set
        e.U, e.Y, e.syn_lr_img, e.syn_lr_txt = engs[0].U, engs[0].Y, engs[0].syn_lr_img, engs[0].syn_lr_txt
    streams = [torch.cuda.Stream() for _ in range(G)]
    g = torch.Generator().manual_seed(0)
    perms = [torch.stack([torch.randperm(100, generator=g) for _ in range(8)]).cuda() for _ in range(8)]
    main = torch.cuda.current_stream()

    def outer(i):
        losses = []
        for j, (e, st) in enumerate(zip(engs, streams)):
            st.wait_stream(main)
            with torch.cuda.stream(st):
                losses.append(e.segment_loss((i * G + j) % 4, ((i * G + j) // 4) % 2, perms[(i * G + j) % 8]))
        for st in streams:
            main.wait_stream(st)
        engs[0].outer_step(sum(losses))
        return losses[0]

    for i in range(3):
        outer(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = outer(3 + i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"G={G} segments in flight: {ms:.3f} ms per outer step = {G * 1e3 / ms:.1f} segment-iterations/s  (loss {float(loss):.6f})", flush=True)

if __name__ == "__main__":
    for G in (1, 2, 3, 4):
        run(G)
