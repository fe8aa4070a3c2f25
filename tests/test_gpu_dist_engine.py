"""Two ranks driving the real engine with a real collective: every rank runs
``DistillEngine.iteration()`` -- its own expert segment, its own minibatch permutations, the CUDA unroll engine, ONE
all-reduce of the packed [dU | dY | out5] buffer, the fused update kernel -- and the result must equal a single process that
loops over the same segments, sums the gradients and applies the same update.

The GPU test tier has one GPU, and NCCL refuses two ranks on one device, so both ranks share cuda:0 and the process
group is gloo (which all-reduces CUDA tensors through the host).  The product code path is identical: it calls
torch.distributed.all_reduce on ws.pack whatever the backend.
"""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

N, B, K, DT, D = 48, 32, 2, 64, 96
ARGS = ["--syn_steps", str(K), "--expert_epochs", "1", "--max_start_epoch", "2", "--num_queries", str(N), "--mini_batch_size", str(B),
        "--lr_img", "10", "--lr_txt", "10", "--lr_lr", "0.01", "--logit_scale_mode", "fork", "--student_dropout", "0.0", "--seed", "5"]
ITERS = 3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    from multimodal_dataset_distillation_b200 import distill
    g = torch.Generator().manual_seed(4)
    U, Y = torch.randn(N, D, generator=g), torch.randn(N, DT, generator=g)
    experts = distill.synthetic_experts(4, 3, DT, D, seed=1)
    return U, Y, experts


def _worker(rank, world, port, outdir, reduce_mode):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multimodal_dataset_distillation_b200 import distill
        args = distill.parse_args(ARGS + ["--grad_reduce", reduce_mode])
        U, Y, experts = _problem()
        eng = distill.DistillEngine(U, Y, experts.cuda(), args, "cuda", rank=rank, world=world)
        segs, losses = [], []
        for _ in range(ITERS):
            e_before = eng.expert_idx
            losses.append(float(eng.iteration()))           # sum of the ranks' losses (out5 rides in the packed buffer)
            segs.append(e_before)
        torch.save(dict(U=eng.U.detach().cpu(), Y=eng.Y.detach().cpu(), lr_img=float(eng.syn_lr_img), lr_txt=float(eng.syn_lr_txt),
                        segs=segs, losses=losses), os.path.join(outdir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("reduce_mode", ["sum", "mean"])
def test_two_ranks_allreduce_equals_single_gpu_sum(tmp_path, reduce_mode):
    from multimodal_dataset_distillation_b200 import distill, ops
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), reduce_mode), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"rank{i}.pt") for i in range(world)]
    # replicas stay identical; ranks worked on different expert trajectories
    assert torch.equal(r[0]["U"], r[1]["U"]) and torch.equal(r[0]["Y"], r[1]["Y"])
    assert r[0]["lr_img"] == r[1]["lr_img"] and r[0]["lr_txt"] == r[1]["lr_txt"]
    assert all(a != b for a, b in zip(r[0]["segs"], r[1]["segs"]))
    assert r[0]["losses"] == r[1]["losses"]
    # single process: loop the same segments with the same per-rank samplers, sum, update
    args = distill.parse_args(ARGS + ["--grad_reduce", reduce_mode])
    U, Y, experts = _problem()
    master = distill.DistillEngine(U, Y, experts.cuda(), args, "cuda", rank=0, world=1)
    samplers = [distill.DistillEngine(U, Y, experts.cuda(), args, "cuda", rank=i, world=world) for i in range(world)]
    scale = (lambda: master.syn_lr_img.detach())
    for it in range(ITERS):
        total = None
        for s in samplers:
            e, ep = s.sample_segment()
            perms = s.draw_perms().cuda()
            res = ops.unrolled_match(experts[e, ep].cuda(), experts[e, ep + 1].cuda(), master.Y.detach(), master.U.detach(),
                                     master.syn_lr_txt.detach(), scale(), perms, None)
            pack = torch.cat([res["dU"].reshape(-1), res["dY"].reshape(-1), res["out5"]])
            total = pack if total is None else total + pack
        master.ws.pack[:total.numel()].copy_(total)
        master.world = world                                   # the mean divides by the number of segments of the step
        master.apply_update(master.ws)
        assert abs(float(total[-3]) - r[0]["losses"][it]) <= 1e-6 * abs(r[0]["losses"][it])
    # gloo sums two values per element: a + b is exact in either order, so the replicas match the loop bit for bit
    torch.testing.assert_close(master.U.detach().cpu(), r[0]["U"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(master.Y.detach().cpu(), r[0]["Y"], rtol=1e-6, atol=1e-7)
    assert abs(float(master.syn_lr_txt) - r[0]["lr_txt"]) < 1e-7 and abs(float(master.syn_lr_img) - r[0]["lr_img"]) < 1e-7
