"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs; gloo in the CPU tests of the host logic).

Only the two places where the path shards naturally use a collective:

* distillation -- every rank runs a different expert segment against the same replicated synthetic set; the
  synthetic-data gradients (dU, dY, dlr_img, dlr_txt) are summed with ONE all-reduce of a packed buffer
  (``allreduce_packed``), after which every rank applies the identical outer update.
* retrieval -- images replicated, captions sharded contiguously.  text->image ranks are final locally (a caption's row
  needs every image, and images are replicated).  image->text needs (1) the best ground-truth caption per image over
  all shards: all-gather of one (score, global index) candidate per image and shard, merged with the same
  "higher score, then lower index" rule the single-GPU kernel uses; (2) the number of captions ranked ahead of that
  candidate: per-shard counts, all-reduce(SUM).  Both exchanges are O(I) integers/floats -- latency-bound.

The device work goes through ``ops`` (CUDA only).  The collective / merge logic is independent of the device and takes
the three local primitives as a ``backend`` object so tests can run it under gloo with a numpy backend.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """torchrun environment -> (rank, world, local_rank); initialises the default process group when world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def bind_to_local_numa(device_index: int) -> list[int] | None:
    """Best effort: restrict this process to the CPUs of the NUMA node its GPU hangs off, so that pinned staging buffers
    allocated AFTERWARDS are first-touched on that node and the per-step host->device copies of 8 ranks do not all cross the
    socket interconnect (the reference keeps expert trajectories in host memory and uploads a segment every iteration,
    distill.py:466-476).  Returns the CPU list it bound to, or None when the topology cannot be read (then nothing changes)."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = []
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except (OSError, ValueError, AttributeError):
        return None


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous balanced shard [lo, hi) of n items; the first n % world shards get one extra item."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_packed(tensors, group=None, op=dist.ReduceOp.SUM):
    """Sum a list of same-dtype tensors across ranks with ONE collective; results are written back in place."""
    if not tensors:
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=op, group=group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
    return tensors


def merge_candidates(scores: torch.Tensor, idx: torch.Tensor):
    """[world, I] per-shard best ground-truth candidates -> global best per image (higher score, then lower index).

    idx == -1 marks "no ground truth in this shard".  Returns (score[I], idx[I]); idx stays -1 if no shard had one.
    """
    valid = idx >= 0
    s = torch.where(valid, scores, torch.full_like(scores, float("-inf")))
    best_s = s.max(dim=0).values
    big = torch.iinfo(idx.dtype).max
    cand = torch.where(valid & (s == best_s.unsqueeze(0)), idx, torch.full_like(idx, big))
    best_i = cand.min(dim=0).values
    none = best_i == big
    best_i = torch.where(none, torch.full_like(best_i, -1), best_i)
    return best_s, best_i


class CudaBackend:
    """Local primitives of the sharded ranking, on the CUDA kernels."""

    def __init__(self):
        from . import ops
        self.ops = ops

    def scores(self, img, txt_shard, scale):
        return self.ops.sim_scores(img, txt_shard, scale, want_t2i=False)[0]        # one GEMM: [I, T_r]

    def best_gt(self, s_i2t, lo, gt_ptr, gt_idx):
        return self.ops.rank_best_gt(s_i2t, lo, gt_ptr, gt_idx)

    def count(self, s_i2t, lo, thr_s, thr_i):
        return self.ops.rank_count(s_i2t, lo, thr_s, thr_i)

    def ranks_t2i(self, s_i2t, txt2img_shard):
        return self.ops.ranks_cols(s_i2t, txt2img_shard)        # a caption's column is complete locally (images replicated)


def sharded_ranks(img, txt_shard, lo: int, txt2img_shard, img2txt_ptr, img2txt_idx, scale: float, group=None,
                  backend=None):
    """Ranks for a caption shard [lo, lo + T_r).  Returns (ranks_i2t[I] -- identical on every rank, ranks_t2i[T_r]).

    img: [I, D] replicated; txt_shard: [T_r, D]; txt2img_shard: int32 [T_r] (image of each local caption);
    img2txt CSR with GLOBAL caption indices.
    """
    be = backend or CudaBackend()
    s_i2t = be.scores(img, txt_shard, scale)
    ranks_t = be.ranks_t2i(s_i2t, txt2img_shard)
    cand_s, cand_i = be.best_gt(s_i2t, lo, img2txt_ptr, img2txt_idx)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        all_s = [torch.empty_like(cand_s) for _ in range(world)]
        all_i = [torch.empty_like(cand_i) for _ in range(world)]
        dist.all_gather(all_s, cand_s, group=group)
        dist.all_gather(all_i, cand_i, group=group)
        thr_s, thr_i = merge_candidates(torch.stack(all_s), torch.stack(all_i))
    else:
        thr_s, thr_i = cand_s, cand_i
    counts = be.count(s_i2t, lo, thr_s, thr_i)
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts, ranks_t


def sharded_ranks_fused(img, txt_shard, lo: int, txt2img_shard, img2txt_ptr, img2txt_idx, scale: float, n_txt_total: int,
                        group=None, shard=None):
    """sharded_ranks without a score matrix: every rank runs the fused tensor-core ranking (ops.FusedRankShard: exact 3xTF32
    candidates, bf16x3-screened count with exact decisions) on its caption shard; the exchange is the same as above -- all-gather
    of one (score, global index) candidate per image and shard, merge, all-reduce of the int32 counts.  Returns
    (ranks_i2t[I] -- identical on every rank, ranks_t2i[T_r]).  Pass a prepared ``shard`` to reuse its workspace across calls."""
    from . import ops
    sh = shard if shard is not None else ops.FusedRankShard(img, txt_shard, lo, txt2img_shard, img2txt_ptr, img2txt_idx, scale)
    cand_s, cand_i = sh.candidates()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        all_s = [torch.empty_like(cand_s) for _ in range(world)]
        all_i = [torch.empty_like(cand_i) for _ in range(world)]
        dist.all_gather(all_s, cand_s, group=group)
        dist.all_gather(all_i, cand_i, group=group)
        thr_s, thr_i = merge_candidates(torch.stack(all_s), torch.stack(all_i))
    else:
        thr_s, thr_i = merge_candidates(cand_s.unsqueeze(0), cand_i.unsqueeze(0))
    counts, ranks_t = sh.count(thr_s, thr_i, 0)
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    ranks_i = torch.where(thr_i >= 0, counts, torch.full_like(counts, int(n_txt_total)))      # no ground truth: never retrieved
    return ranks_i, ranks_t


def sharded_result(ranks_i2t, ranks_t2i_shard, n_txt_total: int, group=None) -> dict:
    """Recall dict of epoch.py:227-244 from sharded ranks (one 3-int all-reduce for the caption side)."""
    ks = torch.tensor([1, 5, 10], device=ranks_i2t.device)
    c_img = (ranks_i2t.unsqueeze(0) < ks.unsqueeze(1)).sum(dim=1)
    c_txt = (ranks_t2i_shard.unsqueeze(0) < ks.unsqueeze(1)).sum(dim=1)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(c_txt, op=dist.ReduceOp.SUM, group=group)
    c_img, c_txt = c_img.cpu().numpy(), c_txt.cpu().numpy()
    n_img = ranks_i2t.numel()
    tr = [100.0 * int(c) / n_img for c in c_img]
    ir = [100.0 * int(c) / n_txt_total for c in c_txt]
    trm, irm = sum(tr) / 3, sum(ir) / 3
    return {"txt_r1": tr[0], "txt_r5": tr[1], "txt_r10": tr[2], "txt_r_mean": trm,
            "img_r1": ir[0], "img_r5": ir[1], "img_r10": ir[2], "img_r_mean": irm, "r_mean": (trm + irm) / 2}
