"""Caption-sharded retrieval over NCCL (BASELINE.json configs[4]): torchrun --nproc-per-node G profiles/retrieval_sharded.py

Every rank holds all images and T/G captions; ranks are compared with the single-GPU result computed on rank 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from multimodal_dataset_distillation_b200 import dist as D, ops

def main():
    rank, world, local = D.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    for I, C, Dm in ((1000, 5, 768), (5000, 5, 768), (25000, 5, 768)):
        T = I * C
        g = torch.Generator(device=dev).manual_seed(0)               # same seed on every rank: replicated data
        img = torch.randn(I, Dm, device=dev, generator=g)
        txt = torch.randn(T, Dm, device=dev, generator=g) + 0.15 * img.repeat_interleave(C, 0)
        img = img / img.norm(dim=1, keepdim=True)
        txt = txt / txt.norm(dim=1, keepdim=True)
        t2i = (torch.arange(T, device=dev) // C).int()
        ptr = (torch.arange(I + 1, device=dev) * C).int()
        idx = torch.arange(T, device=dev).int()
        lo, hi = D.shard_bounds(T, world, rank)
        txt_s, t2i_s = txt[lo:hi].contiguous(), t2i[lo:hi].contiguous()
        for _ in range(2):
            r_i, r_t = D.sharded_ranks(img, txt_s, lo, t2i_s, ptr, idx, 14.285714)
            D.sharded_result(r_i, r_t, T)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r_i, r_t = D.sharded_ranks(img, txt_s, lo, t2i_s, ptr, idx, 14.285714)
        res = D.sharded_result(r_i, r_t, T)
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        full_t = [torch.empty(D.shard_bounds(T, world, r)[1] - D.shard_bounds(T, world, r)[0], dtype=torch.int32, device=dev) for r in range(world)]
        dist.all_gather(full_t, r_t)
        if rank == 0:
            a_i, a_t = ops.sim_rank(img, txt, t2i, ptr, idx, 14.285714)
            ok = bool(torch.equal(a_i, r_i) and torch.equal(a_t, torch.cat(full_t)))
            print(f"G={world} I={I} T={T}: {float(ms):.3f} ms  {I * T / float(ms) / 1e6:.2f} G pairs/s  r_mean={res['r_mean']:.3f}  "
                  f"identical-to-single-GPU={ok}", flush=True)
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
