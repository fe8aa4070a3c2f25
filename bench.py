#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json metric: "distill iters/s (Flickr 100 pairs, syn_steps=8); recall@K eval pairs/s").

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one outer distillation iteration per rank on the Flickr-shaped configuration (BASELINE.json configs[2]):
N = B = 100 synthetic pairs, syn_steps = 8, 768 -> 2304 text_projection head, one expert segment
(theta_start, theta_target) per rank: K-step unroll, matching loss, reverse sweep to the synthetic-data gradients,
gradient all-reduce across ranks (N > 1), momentum-SGD update of the synthetic pairs and the student lr.
`value` = segment-iterations per second summed over all ranks, inputs already resident in HBM.
`e2e`   = the same iteration through the public Python API with the segment coming from pinned HOST memory
          (H2D of theta_start/theta_target/perms inside the timed region, loss read back D2H every step).
The retrieval metric (pairs/s, configs[0] shape 1000 x 5000 x 768) is reported on the same line under "retrieval".

`--impl reference` times the reference's own mechanism on the host CPU: the oracle restatement of
distill.py:509-606 (torch autograd, create_graph=True double backward, all host threads).  distill.py itself is not
importable (clip/timm/kornia/BERT download), see oracle/distill_ref.py.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: keep NCCL's "NCCL version ..." banner (printed to stdout at VERSION level) out of it
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

CFG = dict(N=100, B=100, K=8, dt=768, d=2304, experts=4, snapshots=3)
RET = dict(I=1000, C=5, D=768)
METRIC = "distill iters/s (Flickr 100 pairs, syn_steps=8)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf=p["bf16_tflops_sustained"], tf_burst=p.get("bf16_tflops", p["bf16_tflops_sustained"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf=1590.0, tf_burst=1590.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.stop_flag, self.index = [], False, index
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop_flag = True
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def make_experts(seed):
    from multimodal_dataset_distillation_b200 import distill
    return distill.synthetic_experts(CFG["experts"], CFG["snapshots"], CFG["dt"], CFG["d"], seed=seed, step=0.01)


def make_pairs(seed=0):
    g = torch.Generator().manual_seed(seed)
    U = torch.randn(CFG["N"], CFG["d"], generator=g)
    Y = torch.randn(CFG["N"], CFG["dt"], generator=g) * 0.5253 - 0.0094
    return U, Y


def bench_args():
    from multimodal_dataset_distillation_b200 import distill
    return distill.parse_args(["--syn_steps", str(CFG["K"]), "--expert_epochs", "1", "--max_start_epoch", "2",
                               "--num_queries", str(CFG["N"]), "--mini_batch_size", str(CFG["B"]), "--lr_img", "1000",
                               "--lr_txt", "1000", "--lr_lr", "0.01", "--logit_scale_mode", "upstream",
                               "--student_dropout", "0.1"])


def algorithmic_bytes_per_iteration(K, P):
    # forward unroll reads theta_k and writes theta_{k+1} (2K passes), matching loss reads 3
    # vectors, reverse sweep reads theta_k, a_{k+1}, writes a_k plus one extra weight pass (4K)
    return 4 * P * (2 * K + 3 + 4 * K)


def dominant_kernel_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed `ncu --set full`
    capture of the CURRENT kernel (profiles/ncu_dominant.json, written by profiles/summarise_ncu.py from the .ncu-rep);
    None when no capture of this round's kernel is committed."""
    path = os.path.join(ROOT, "profiles", "ncu_dominant.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        d = json.load(f)
    return d.get("dram_bytes_read", 0) + d.get("dram_bytes_write", 0), d


def bench_dominant_kernel(dev, experts, reps=20):
    """Time the dominant kernel alone: tc_gemm_kernel<K-major,K-major,3xTF32,EpiPartial> as launched for f = h W2^T
    (M=100 activations x N=K=2304 weights) with CUDA events on its stream.  The weight operand rotates over the
    resident expert snapshots (12 x 21 MB > 126 MB L2), so every launch streams its weights from HBM."""
    import ctypes as C
    from multimodal_dataset_distillation_b200._lib import lib, check
    M, N, K = CFG["B"], CFG["d"], CFG["d"]
    o_w2 = CFG["d"] * CFG["dt"] + CFG["d"]
    flat = experts.reshape(-1, experts.shape[-1])
    ws_bytes = lib().vldd_bench_skinny_gemm_workspace_bytes(M, N, K)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    A = torch.randn(M, K, device=dev)
    splits = C.c_int(0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def launch(i):
        W = flat[i % flat.shape[0], o_w2:o_w2 + N * K]
        check(lib().vldd_bench_skinny_gemm(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), M, N, K,
                                           C.c_void_p(ws.data_ptr()), ws_bytes, C.byref(splits), st), "bench_skinny_gemm")
    for i in range(flat.shape[0]):
        launch(i)                      # warm-up: tensor maps encoded, kernel loaded
    torch.cuda.synchronize()
    # `group` launches back to back per event pair (each on a different resident snapshot), so that the event / launch
    # latency of a lone 10-us kernel is amortised the way it is inside the iteration's CUDA graph
    group = flat.shape[0]
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for r, (e0, e1) in enumerate(evs):
        e0.record()
        for i in range(group):
            launch(r * group + i)
        e1.record()
    torch.cuda.synchronize()
    us = sorted(1e3 * e0.elapsed_time(e1) / group for e0, e1 in evs)
    avg_us = sum(us) / len(us)
    # algorithmic bytes per launch: weights once + activations once + the fp32 result once (4*P per weight pass)
    abytes = 4 * (N * K + M * K + M * N)
    return dict(avg_us=avg_us, median_us=us[len(us) // 2], abytes=abytes, splits=splits.value,
                flops=2 * M * N * K, launches=reps * group)


def gpu_eager_baseline(dev, max_seconds=20.0, max_iters=10):
    """The reference's own mechanism on the SAME B200 through PyTorch eager (BASELINE.md section 3 "B3", SURVEY 8d: the
    honest bar): oracle restatement of distill.py:509-606 -- autograd.grad(create_graph=True) per step, backward()
    through the unroll -- with every tensor on the GPU, fp32 matmuls (torch default: TF32 off)."""
    from oracle import distill_ref as R
    pr = R.make_problem(N=CFG["N"], B=CFG["B"], K=CFG["K"], dt=CFG["dt"], d=CFG["d"], seed=0, dropout=True)
    pr = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in pr.items()}
    for _ in range(2):
        R.unrolled_match_autograd(**pr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.perf_counter()
    e0.record()
    while n < max_iters and (time.perf_counter() - t0) < max_seconds:
        R.unrolled_match_autograd(**pr)
        torch.cuda.synchronize()
        n += 1
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return {"value": 1e3 / ms, "unit": "iters/s", "ms_per_step": ms, "iterations": n, "kind": "port",
            "what": "PyTorch eager on this GPU: oracle/distill_ref.py::unrolled_match_autograd (torch %s, fp32, "
                    "allow_tf32=%s), same Flickr-shaped workload, inputs resident" % (torch.__version__,
                                                                                      torch.backends.cuda.matmul.allow_tf32)}


def _time_rotating(fn, n_variants, reps=5):
    """Average us per call of fn(i) over `reps` rounds of `n_variants` calls, each on different buffers (> L2 in total)."""
    for i in range(n_variants):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        for i in range(n_variants):
            fn(i)
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (reps * n_variants)


def bench_streaming_kernels(dev, experts):
    """Achieved HBM GB/s of the streaming kernels (north_star: "achieved HBM GB/s for the streaming kernels ... against
    B200 peak"; distill.py:582-598, 611-613; epoch.py:219-244), each launched back to back on rotating buffers whose
    total exceeds the 126 MB L2 (12 resident snapshots of 28 MB + 6 scratch vectors), CUDA events on the launch stream."""
    import ctypes as C
    from multimodal_dataset_distillation_b200 import ops
    from multimodal_dataset_distillation_b200._lib import lib, check
    pk = peaks()
    flat = experts.reshape(-1, experts.shape[-1])
    S, P = flat.shape
    outs = [torch.empty(P, device=dev) for _ in range(6)]
    bufs = [torch.zeros(P, device=dev) for _ in range(6)]
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ptr = lambda t: C.c_void_p(t.data_ptr())
    res = {}

    def entry(name, us, nbytes, per_elem, site):
        gbs = nbytes / (us * 1e-6) / 1e9
        res[name] = {"avg_launch_us": us, "algorithmic_bytes_per_launch": nbytes, "bytes_per_element": per_elem,
                     "achieved_gbs": gbs, "frac": gbs / pk["hbm"], "reference_site": site}

    # operand sets are disjoint from launch to launch: a buffer comes back only after > 300 MB of other traffic (L2 = 126 MB)
    lr = torch.full((1,), 0.1, device=dev)
    us = _time_rotating(lambda i: check(lib().vldd_flat_sgd_step(ptr(flat[(2 * i) % S]), ptr(flat[(2 * i + 1) % S]), ptr(lr), ptr(outs[i % 6]), P, st()), "sgd"), 6)
    entry("flat_sgd_step", us, 12 * P, 12, "distill.py:582-583")
    out3 = torch.empty(3, device=dev)
    scratch = torch.zeros(lib().vldd_match_loss_scratch_bytes(), dtype=torch.uint8, device=dev)
    us = _time_rotating(lambda i: check(lib().vldd_match_loss_fwd(ptr(flat[(3 * i) % S]), ptr(flat[(3 * i + 1) % S]), ptr(flat[(3 * i + 2) % S]), P, ptr(out3), ptr(scratch), st()), "ml"), 4)
    entry("match_loss_fwd", us, 12 * P, 12, "distill.py:588-598 (stand-alone form: 3 reads, ticketed finish inside the kernel)")
    us = _time_rotating(lambda i: check(lib().vldd_match_loss_bwd(ptr(flat[(2 * i) % S]), ptr(flat[(2 * i + 1) % S]), ptr(out3), None, ptr(outs[i % 6]), P, st()), "mlb"), 6)
    entry("match_loss_bwd", us, 12 * P, 12, "distill.py:606 (d/d theta_K of the ratio; stand-alone form)")
    den = torch.ones(1, device=dev)
    us = _time_rotating(lambda i: check(lib().vldd_match_final(ptr(flat[(2 * i) % S]), ptr(flat[(2 * i + 1) % S]), ptr(den), P, None, ptr(outs[i % 6]), ptr(scratch), st()), "mlf"), 6)
    entry("match_final_pass", us, 12 * P, 12, "distill.py:588-598 + 606: what the engine runs -- numerator partials + adjoint in one pass "
          "(2 reads + 1 write); the denominator comes from the staging pass, the partials are added by the call's last kernel")
    us = _time_rotating(lambda i: check(lib().vldd_momentum_sgd(ptr(outs[i % 6]), ptr(flat[i % S]), ptr(bufs[i % 6]), 1e-9, 0.5, 0, P, st()), "mom"), 12)
    entry("momentum_sgd", us, 20 * P, 20, "distill.py:233-241, 611-613")
    # rank kernels on a score matrix larger than L2 (COCO eval shape: 5000 x 25000 fp32 = 500 MB, read once per direction)
    I, Cc = 5000, 5
    T = I * Cc
    Smat = torch.randn(I, T, device=dev)
    t2i = (torch.arange(T, device=dev, dtype=torch.int32) // Cc).contiguous()
    gptr = (torch.arange(I + 1, device=dev, dtype=torch.int32) * Cc).contiguous()
    gidx = torch.arange(T, device=dev, dtype=torch.int32)
    us = _time_rotating(lambda i: ops.ranks_from_scores(Smat, None, t2i, gptr, gidx), 1, reps=5)
    entry("ranks_rows", us, 4 * I * T, 4, "epoch.py:227-233 (image->text)")
    us = _time_rotating(lambda i: ops.ranks_cols(Smat, t2i), 1, reps=5)
    entry("ranks_cols", us, 4 * I * T, 4, "epoch.py:236-241 (text->image, read column-wise from the same matrix)")
    return res


def gpu_retrieval_set(I, Cc, D, dev, seed=0):
    """Synthetic retrieval set generated on the device (identical on every rank: same seed, same GPU model):
    SURVEY 8d config 1 recipe -- N(0,1) embeddings, captions pulled 0.15 towards their image, rows normalised."""
    g = torch.Generator(device=dev).manual_seed(seed)
    img = torch.randn(I, D, generator=g, device=dev)
    txt = torch.randn(I * Cc, D, generator=g, device=dev) + 0.15 * img.repeat_interleave(Cc, dim=0)
    img = img / img.norm(dim=1, keepdim=True)
    txt = txt / txt.norm(dim=1, keepdim=True)
    T = I * Cc
    t2i = (torch.arange(T, device=dev, dtype=torch.int32) // Cc).contiguous()
    gptr = (torch.arange(I + 1, device=dev, dtype=torch.int32) * Cc).contiguous()
    gidx = torch.arange(T, device=dev, dtype=torch.int32)
    return img.contiguous(), txt.contiguous(), t2i, gptr, gidx


def bench_retrieval_large(dev, world, rank, reps=3):
    """configs[3]/[4]: 5000 x 25000 and 25000 x 125000 (D = 768).  One GPU: vldd_sim_rank_fused.  N > 1: captions sharded
    per rank (dist.sharded_ranks_fused: the same fused ranking per [I, T/N] shard -- no score matrix --, two 8 B/image
    all-gathers + one int32 all-reduce), ranks asserted equal to rank 0's single-GPU result; time = max over ranks."""
    import torch.distributed as dist
    from multimodal_dataset_distillation_b200 import ops, dist as D
    pk = peaks()
    out = []
    for I in (5000, 25000):
        Cc, Dm = 5, 768
        T = I * Cc
        img, txt, t2i, gptr, gidx = gpu_retrieval_set(I, Cc, Dm, dev, seed=I)
        def single():
            return ops.sim_rank_fused(img, txt, t2i, gptr, gidx, 14.285714)
        entry = {"workload": f"{I} images x {T} captions, {Dm}-d", "pairs": I * T}
        if world == 1:
            single()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                r1, r2 = single()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            entry.update(path="vldd_sim_rank_fused (one GPU, no score matrix in HBM)")
        else:
            lo, hi = D.shard_bounds(T, world, rank)
            txt_s, t2i_s = txt[lo:hi].contiguous(), t2i[lo:hi].contiguous()
            shard = ops.FusedRankShard(img, txt_s, lo, t2i_s, gptr, gidx, 14.285714)
            def sharded():
                return D.sharded_ranks_fused(img, txt_s, lo, t2i_s, gptr, gidx, 14.285714, T, shard=shard)
            sharded()
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                ri, rt = sharded()
            e1.record()
            dist.barrier(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
            # parity: the sharded ranks equal the single-GPU ranks (rank 0 computes those)
            sizes = [D.shard_bounds(T, world, r)[1] - D.shard_bounds(T, world, r)[0] for r in range(world)]
            parts = [torch.empty(n, dtype=torch.int32, device=dev) for n in sizes]
            dist.all_gather(parts, rt.contiguous())
            ok = None
            if rank == 0:
                r1, r2 = single()
                ok = bool(torch.equal(r1, ri)) and bool(torch.equal(r2, torch.cat(parts)))
                assert ok, "sharded ranks differ from the single-GPU ranks"
            entry.update(path=f"dist.sharded_ranks_fused: captions sharded {world}-way, images replicated, fused ranking per shard",
                         collectives="2 x all_gather of 8 B/image (best ground-truth candidate per shard) + 1 x int32 "
                                     "all_reduce of I counts (NCCL over NVLink); latency-bound, the GEMM dominates",
                         ranks_equal_single_gpu=ok)
            del txt_s, t2i_s, shard
        tflops = 2.0 * I * T * Dm / (ms / 1e3) / 1e12
        entry.update(ms=ms, value=I * T / (ms / 1e3), unit="pairs/s",
                     roofline={"bound": "tensor", "achieved": tflops, "peak": pk["tf"], "unit": "TFLOP/s", "frac": tflops / pk["tf"],
                               "note": "useful 2*I*T*D flops counted once for both directions against sustained dense bf16",
                               # what the tensor cores actually execute: the screen runs THREE bf16 products per pair (plus the
                               # exact passes, not counted) -- the whole call's time against that work and the burst peak
                               "mma_work_tflops": 3.0 * tflops, "mma_work_frac_of_burst_peak": 3.0 * tflops / (world * pk["tf_burst"]),
                               "mma_work_note": "3 bf16 MMAs per pair (screen) over the WHOLE call's time, all GPUs; ncu on the "
                                                "screen kernel alone: 91 % tensor-pipe activity (profiles/ncu_r02z_retrieval_25k.txt)"})
        out.append(entry)
        del img, txt
        torch.cuda.empty_cache()
    return out


def allreduce_parity_check(eng, dev, world, rank):
    """N > 1: the all-reduced (dU, dY, dlr, dscale) of N segments (one per rank) equals the sum a single GPU gets by
    looping the same N segments (SURVEY section 4 test plan).  Dropout off so that every rank's call is reproducible."""
    import torch.distributed as dist
    from multimodal_dataset_distillation_b200 import ops, dist as D
    K, B, N = CFG["K"], CFG["B"], CFG["N"]

    def segment(r):
        ex = make_experts(100 + r)[0].to(dev)                     # expert 0 of rank r's trajectories
        g = torch.Generator().manual_seed(1000 + r)
        perms = torch.stack([torch.randperm(N, generator=g)[:B] for _ in range(K)]).to(dev)
        res = ops.unrolled_match(ex[0], ex[1], eng.Y.detach(), eng.U.detach(), eng.syn_lr_txt.detach(), eng.fixed_scale, perms, None)
        return [res["dU"].clone(), res["dY"].clone(), res["out5"][3:5].clone()]

    mine = segment(rank)
    D.allreduce_packed(mine)
    err = None
    if rank == 0:
        tot = None
        for r in range(world):
            part = segment(r)
            tot = part if tot is None else [a + b for a, b in zip(tot, part)]
        err = max(float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)) for a, b in zip(mine, tot))
        assert err < 1e-5, f"all-reduced gradients differ from the single-GPU sum: {err}"
    dist.barrier()
    return {"segments": world, "max_rel_err_vs_single_gpu_sum": err,
            "what": "NCCL all-reduce of packed [dU | dY | dlr, dscale] from one segment per rank vs rank 0 looping the same "
                    "segments and summing (fp32; summation order differs)"}


def workload_config(world):
    return {"workload": "configs[2]: full distill inner loop, Flickr30K-shaped (N=B=100 pairs, syn_steps=8, "
                        "expert_epochs=1, max_start_epoch=2, text_projection 768->2304 in train mode (fresh "
                        "dropout-0.1 masks every iteration), image side = frozen 2304-d embeddings)",
            "unit_of_work": "one expert segment: 8-step unroll + matching loss + reverse sweep + outer SGD; "
                            "one segment per rank per step, grads all-reduced (NCCL) when n_gpus > 1",
            "l2_policy": "inputs larger than L2: steps rotate over 4 experts x 2 start epochs (340 MB of "
                         "snapshots) and the per-iteration working set is ~450 MB vs 126 MB L2",
            "parallelism": f"dp{world} (one expert segment per GPU)"}


def run_ours(opt):
    import torch.distributed as dist
    from multimodal_dataset_distillation_b200 import distill, ops, epoch
    from multimodal_dataset_distillation_b200._lib import lib
    lib()                                            # fail loudly here if the CUDA library is missing
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = None
    if world > 1:
        from multimodal_dataset_distillation_b200 import dist as dist_mod
        numa_cpus = dist_mod.bind_to_local_numa(local)      # pinned staging buffers of this rank on its GPU's NUMA node
        dist.init_process_group("nccl", device_id=dev)
    args = bench_args()
    U, Y = make_pairs(0)                             # replicated synthetic set
    experts_host = make_experts(100 + rank)          # each rank owns different expert trajectories
    experts = experts_host.to(dev)
    eng = distill.DistillEngine(U, Y, experts, args, dev)
    K, B, N = CFG["K"], CFG["B"], CFG["N"]
    g = torch.Generator().manual_seed(rank)
    perm_sets = [torch.stack([torch.randperm(N, generator=g)[:B] for _ in range(K)]).to(dev) for _ in range(8)]

    def step(i):
        e, s = i % CFG["experts"], (i // CFG["experts"]) % 2          # rotate segments: inputs (4 x 3 x 28 MB) exceed L2
        return eng.step_fast(e, s, perm_sets[i % 8])                   # engine call (+ all-reduce) + fused update

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(opt.warmup):
        step(i)
    sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(local)                       # sampled across both timed regions (device-resident and end-to-end)
    clk.__enter__()
    sync()
    launches0 = lib().vldd_kernel_launch_count()
    ev0.record()
    for i in range(opt.steps):
        step(opt.warmup + i)
    ev1.record()
    sync()
    gpu_launches = int(lib().vldd_kernel_launch_count() - launches0)     # counted by the library's launcher, not a formula
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    ms_per_step = ms / opt.steps
    value = world * opt.steps / (ms / 1e3)

    if opt.profile:                                  # ncu target: nothing but the loop above
        clk.__exit__(None, None, None)
        if world > 1:
            dist.destroy_process_group()
        if rank != 0:
            return None
        return {"metric": METRIC, "value": value, "unit": "iters/s", "n_gpus": world, "steps": opt.steps,
                "warmup": opt.warmup, "ms_per_step": ms_per_step, "gpu_launches": gpu_launches,
                "profile_only": "developer run (--profile): no end-to-end, roofline, retrieval or baseline legs -- not a bench line"}

    # ---- end-to-end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region ----
    # Expert trajectories stay in host memory as in the reference (distill.py:466-476 uploads the segment every
    # iteration).  "streamed": distill.SegmentPrefetcher copies segment i+1 on a copy stream while segment i is processed
    # (every step moves 2 snapshots = 56.7 MB).  "cached": distill.SegmentCache keeps the most recently used snapshots in a
    # device-side LRU and uploads only misses (the bench's 4 experts x 3 snapshots all fit, so steady state moves only
    # the minibatch permutations).  Both read the loss back to the host every step.
    P = ops.head_numel(CFG["dt"], CFG["d"])
    perms_host = [p.cpu().pin_memory() for p in perm_sets]
    seg = lambda i: (i % CFG["experts"], (i // CFG["experts"]) % 2)

    def run_e2e(pre, late_read=True):
        pre.prefetch(*seg(0), 1)
        # The loss of EVERY step is read on the host; with late_read one step late: step i+1 is enqueued before the host blocks
        # on step i's 4-byte result, so the host's launch work (~60 us) overlaps the GPU instead of following it.  (A non-finite
        # loss is still kept from the parameters in time: vldd_outer_update checks it on the device.)  distill.StepIO keeps
        # the small per-step copies (indices up, loss down) out of the compute stream, where they would queue behind the
        # segment upload on the copy engine and stall the kernels behind them.
        io = distill.StepIO(K, B, dev, pre.copy_stream)
        io.upload_perms(0, perms_host[0])
        seen = []

        def e2e_step(i):
            sl = pre.get()
            loss = eng.step_fast(perms=io.perms_for(i), theta0=sl["th0"], theta_tgt=sl["tgt"])
            pre.release(sl)
            io.step_done(i, loss)
            pre.prefetch(*seg(i + 1), 1)                                     # next segment's H2D overlaps this iteration
            io.upload_perms(i + 1, perms_host[(i + 1) % 8])
            j = i - 1 if late_read else i
            if j >= 0:
                seen.append(io.loss(j))                                      # the user reads every iteration's loss

        n_warm = max(opt.warmup, 8)                                          # one full rotation of the 8 segments
        for i in range(n_warm):
            e2e_step(i)
        sync()
        bytes0 = getattr(pre, "h2d_bytes", None)
        t0 = time.perf_counter()
        for i in range(opt.steps):
            e2e_step(n_warm + i)
        seen.append(io.loss(n_warm + opt.steps - 1))
        sync()
        e2e_s = time.perf_counter() - t0
        assert all(v == v for v in seen[-opt.steps:]), "NaN loss in the end-to-end loop"
        te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        copied = None if bytes0 is None else (pre.h2d_bytes - bytes0) / opt.steps
        return world * opt.steps / float(te), copied

    # streamed: 57 MB of H2D per step.  Headline: every step's loss read one step late (the host enqueues step i+1 while the GPU
    # works on step i); also reported: the loss read in lock step (the host's launch work then follows each step).
    e2e_lock, _ = run_e2e(distill.SegmentPrefetcher(experts_host, dev), late_read=False)
    e2e_value, _ = run_e2e(distill.SegmentPrefetcher(experts_host, dev), late_read=True)
    e2e_cached, cached_bytes = run_e2e(distill.SegmentCache(experts_host, dev, capacity=16))
    clk.__exit__(None, None, None)
    h2d = 2 * P * 4 + K * B * 8
    d2h = 4

    # ---- throughput mode (extra, N = 1 only): two segments of one outer step in flight on one GPU ----
    conc = None
    if world == 1:
        def conc_step(i):
            segs = [((2 * i + j) % CFG["experts"], ((2 * i + j) // CFG["experts"]) % 2) for j in range(2)]
            return eng.segments_step(segs, [perm_sets[(2 * i + j) % 8] for j in range(2)])
        for i in range(3):
            conc_step(i)
        sync()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        n_outer = max(opt.steps // 2, 1)
        for i in range(n_outer):
            conc_step(3 + i)
        c1.record()
        sync()
        conc = {"segments_in_flight": 2, "value": 2 * n_outer / (c0.elapsed_time(c1) / 1e3), "unit": "segment-iterations/s",
                "ms_per_outer_step": c0.elapsed_time(c1) / n_outer,
                "note": "NOT the headline: two expert segments per outer step (the G-segment minibatch of the multi-GPU run, "
                        "on one GPU), each on its own stream and workspace, so that one segment's pipeline fills and drains "
                        "overlap the other's work (DistillEngine.segments_step)"}

    out = None
    if rank == 0:
        pk = peaks()
        abytes = algorithmic_bytes_per_iteration(K, P)
        achieved = abytes / (ms_per_step / 1e3) / 1e9
        dk = bench_dominant_kernel(dev, experts)
        dk_achieved = dk["abytes"] / (dk["avg_us"] * 1e-6) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": "iters/s", "n_gpus": world, "steps": opt.steps, "warmup": opt.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(world),
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": "iters/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "numa_local_staging_cpus": (len(numa_cpus) if numa_cpus else None),
                    "streamed_with_lock_step_loss_read": e2e_lock,
                    "note": "streamed: segment (theta_start, theta_target) and minibatch indices copied from pinned host memory "
                            "every step (prefetched one iteration ahead on a copy stream); every step's loss is read back on the "
                            "host, one step late so that the host's launch work overlaps the GPU; the small per-step copies "
                            "stay out of the compute stream (distill.StepIO), where they queued behind the segment upload",
                    "cached": {"value": e2e_cached, "unit": "iters/s", "h2d_bytes_per_step": (cached_bytes or 0) + K * B * 8,
                               "d2h_bytes_per_step": d2h,
                               "note": "same host-resident trajectories behind distill.SegmentCache (device-side LRU of uploaded "
                                       "snapshots, 16 slots): only misses are copied; the bench's 12 snapshots all fit"}},
            "gpu_launches": gpu_launches,
            "roofline": {"bound": "hbm", "achieved": dk_achieved, "peak": pk["hbm"], "unit": "GB/s",
                         "frac": dk_achieved / pk["hbm"], "traffic": dominant_kernel_traffic()[0],
                         "traffic_source": (dominant_kernel_traffic()[1] or {}).get("source"),
                         "peak_source": pk["src"],
                         "kernel": "vldd::tc::tc_gemm_kernel<K-major,K-major,3xTF32,EpiPartialTma,BN=128> launched as "
                                   "f = h W2^T (M=100, N=K=2304, split-K %d, slabs written by TMA stores): the GEMM family is "
                                   "~65%% of the iteration's serialised kernel time (profiles/launches_r02z_summary.txt)"
                                   % dk["splits"],
                         "algorithmic_bytes_per_launch": dk["abytes"], "avg_launch_us": dk["avg_us"],
                         "median_launch_us": dk["median_us"], "launches_timed": dk["launches"],
                         "tensor_tflops_3xtf32_equiv": 3 * dk["flops"] / (dk["avg_us"] * 1e-6) / 1e12,
                         "whole_iteration": {"achieved": achieved, "frac": achieved / pk["hbm"], "unit": "GB/s",
                                             "algorithmic_bytes_per_step": abytes,
                                             "definition": "4*P*(2K+3+4K) bytes per iteration / ms_per_step"}},
        }
        if conc is not None:
            out["concurrent_segments"] = conc
        out["retrieval"] = bench_retrieval(dev, opt)
        if world == 1:                                   # the baselines and per-kernel legs belong to the N = 1 line
            out["kernels"] = bench_streaming_kernels(dev, experts)
            out["gpu_eager_baseline"] = gpu_eager_baseline(dev)
            out["cpu_baseline"] = cpu_baseline_distill(max_seconds=20.0)
    # every rank takes part in the sharded retrieval sweep and the all-reduce parity check
    if world > 1:
        parity = allreduce_parity_check(eng, dev, world, rank)
    large = bench_retrieval_large(dev, world, rank)
    if rank == 0:
        out["retrieval"]["large"] = large
        if world > 1:
            out["parity_check"] = parity
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def bench_retrieval(dev, opt):
    """Secondary metric: recall@K eval pairs/s at Flickr test shape (configs[0])."""
    from multimodal_dataset_distillation_b200 import ops
    from oracle import retrieval_ref as RR
    I, C, D = RET["I"], RET["C"], RET["D"]
    T = I * C
    img, txt = RR.synthetic_retrieval(I, C, D, seed=0)
    txt2img, img2txt = RR.flickr_maps(I, C)
    t2i, ptr, idx = ops.maps_to_arrays(txt2img, img2txt, I, T)
    d = lambda a: torch.from_numpy(a).to(dev)
    img_d, txt_d, t2i_d, ptr_d, idx_d = d(img), d(txt), d(t2i), d(ptr), d(idx)
    ws = torch.empty(ops.lib().vldd_sim_rank_workspace_bytes(I, T, D), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    reps = max(5, opt.steps)
    for _ in range(3):
        ops.sim_rank(img_d, txt_d, t2i_d, ptr_d, idx_d, 14.285714, ws)
    tot = 0.0
    for _ in range(reps):
        flush.zero_()                                   # L2 flush between timed iterations (inputs are < L2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r1, r2 = ops.sim_rank(img_d, txt_d, t2i_d, ptr_d, idx_d, 14.285714, ws)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / reps
    # e2e: the reference's itm_eval signature with HOST score matrices (pinned), ranks + recall back on the host
    S = (np.float32(14.285714) * img) @ txt.T
    s1 = torch.from_numpy(S).pin_memory()
    s2 = torch.from_numpy(np.ascontiguousarray(S.T)).pin_memory()
    for _ in range(2):
        ops.itm_eval_host(s1.numpy(), s2.numpy(), t2i, ptr, idx)
    t0 = time.perf_counter()
    for _ in range(reps):
        res = ops.itm_eval_host(s1.numpy(), s2.numpy(), t2i, ptr, idx)
    e2e_s = (time.perf_counter() - t0) / reps
    # CPU baseline: the oracle restatement of itm_eval on the same matrices (single thread, numpy)
    t0 = time.perf_counter()
    ref = RR.recall_dict(RR.ranks_vectorised(S, ptr, idx), RR.ranks_vectorised(np.ascontiguousarray(S.T), np.arange(T + 1, dtype=np.int32), t2i))
    cpu_s = time.perf_counter() - t0
    pk = peaks()
    tflops = 2.0 * I * T * D / (ms / 1e3) / 1e12
    return {"metric": "recall@K eval pairs/s", "workload": f"configs[0]: {I} images x {T} captions, {D}-d",
            "value": I * T / (ms / 1e3), "unit": "pairs/s", "ms": ms,
            "roofline": {"bound": "tensor", "achieved": tflops, "peak": pk["tf"], "unit": "TFLOP/s", "frac": tflops / pk["tf"],
                         "peak_source": pk["src"] + " (sustained dense bf16; the kernel runs 3 tf32 MMAs per fp32 product, "
                                                    "so its ceiling is peak/6)",
                         "note": "2*I*T*D useful flops counted once for both directions; at this size (7.7 GFLOP) the "
                                 "call is launch-latency-bound, see profiles/retrieval_sweep_r01.txt for the large shapes"},
            "what": "embeddings resident in HBM -> similarity GEMM -> ranks of both directions (vldd_sim_rank)",
            "e2e": {"value": I * T / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": 2 * I * T * 4 + (T + I + 1 + T) * 4,
                    "d2h_bytes_per_step": 32, "what": "itm_eval(host score matrices) through vldd_itm_eval_host"},
            "cpu_baseline": {"value": I * T / cpu_s, "unit": "pairs/s", "cores": 1, "kind": "port",
                             "sample": "oracle itm_eval restatement (numpy) on the same 1000x5000 matrices, ranking only"},
            "r_mean": res["r_mean"], "r_mean_cpu": ref["r_mean"]}


def cpu_baseline_distill(max_seconds=20.0, max_iters=8):
    """Oracle (torch CPU restatement of distill.py:509-606, autograd double backward) on the host cores."""
    from oracle import distill_ref as R
    torch.set_num_threads(host_threads())
    pr = R.make_problem(N=CFG["N"], B=CFG["B"], K=CFG["K"], dt=CFG["dt"], d=CFG["d"], seed=0, dropout=True)
    R.unrolled_match_autograd(**pr)                       # warm-up (thread pools, allocator)
    n, t0 = 0, time.perf_counter()
    while n < max_iters and (time.perf_counter() - t0) < max_seconds:
        R.unrolled_match_autograd(**pr)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "iters/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} full iterations of the same Flickr-shaped workload (oracle/distill_ref.py::unrolled_match_autograd, "
                      f"torch {torch.__version__} CPU, {torch.get_num_threads()} threads of {os.cpu_count()} cpus)"}


def host_threads():
    """Threads for the CPU arms.  torchrun exports OMP_NUM_THREADS=1 to every rank, which would make the N > 1 reference
    run single-threaded; the reference arm runs on rank 0 alone, so it gets the whole box either way: one thread per
    physical core (torch's own default when nothing is forced: half the logical CPUs)."""
    n = os.cpu_count() or 2
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    return max(1, n // 2)


def run_reference(opt):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    torch.set_num_threads(host_threads())
    from oracle import distill_ref as R
    pr = R.make_problem(N=CFG["N"], B=CFG["B"], K=CFG["K"], dt=CFG["dt"], d=CFG["d"], seed=0, dropout=True)
    steps, warm = opt.steps, opt.warmup
    for _ in range(max(warm, 1)):
        R.unrolled_match_autograd(**pr)
    t0 = time.perf_counter()
    for _ in range(steps):
        R.unrolled_match_autograd(**pr)
    dt = time.perf_counter() - t0
    v = steps / dt
    sample = (f"{steps} full iterations (one step = one whole iteration, ~0.15 s) of the Flickr-shaped workload on the host CPU, "
              f"{torch.get_num_threads()} threads of {os.cpu_count()} logical cpus: oracle restatement of distill.py:509-606 "
              f"with torch autograd double backward")
    return {"impl": "reference", "metric": METRIC, "value": v, "unit": "iters/s", "n_gpus": opt.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(1),
            "cpu_baseline": {"value": v, "unit": "iters/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--profile", action="store_true",
                    help="developer aid for ncu: only the device-resident distill loop (value / ms_per_step / gpu_launches), none "
                         "of the other legs (end to end, per-kernel, retrieval, baselines); the JSON line says so")
    opt = ap.parse_args()
    opt.warmup = max(opt.warmup, 3) if opt.impl == "ours" else opt.warmup
    # stdout must carry exactly ONE JSON line: libraries write banners to file descriptor 1 behind Python's back (NCCL prints
    # "NCCL version ..." when its first communicator comes up), so everything but the final line goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        out = run_reference(opt) if opt.impl == "reference" else run_ours(opt)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
