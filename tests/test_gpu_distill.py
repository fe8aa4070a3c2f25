"""CUDA head / InfoNCE / unroll engine vs the oracle and the reference-generated goldens, through the C ABI.

Tolerance (BASELINE.json north_star): losses and synthetic-data gradients within 1e-4 relative in fp32.
For tensors "relative" is norm-wise: max|got - ref| <= 1e-4 * max|ref|.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import distill_ref as R

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def rel_err(got, ref):
    got, ref = torch.as_tensor(got).double().cpu(), torch.as_tensor(ref).double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-300))


def to_cuda(pr):
    return {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in pr.items()}


@pytest.mark.parametrize("rows,dt,d,drop", [(5, 12, 20, False), (7, 10, 16, True), (100, 768, 2304, False), (130, 768, 2304, True),
                                            (1, 768, 2304, False)])
def test_proj_head_forward(rows, dt, d, drop):
    from multimodal_dataset_distillation_b200 import ops
    pr = R.make_problem(N=rows, B=rows, K=1, dt=dt, d=d, seed=11, dropout=drop)
    mask = pr["masks"][0] if drop else None
    ref = R.head_forward(pr["theta0"].double(), pr["Y"].double(), dt, d, None if mask is None else mask.double())
    got = ops.proj_head_forward(pr["theta0"].cuda(), pr["Y"].cuda(), d, None if mask is None else mask.cuda())
    assert rel_err(got, ref) < RTOL
    gotn = ops.proj_head_forward(pr["theta0"].cuda(), pr["Y"].cuda(), d, None if mask is None else mask.cuda(), normalise=True)
    assert rel_err(gotn, R.row_normalise(ref)) < RTOL


def test_proj_head_forward_matches_reparam_golden():
    """ReparamModule(ProjectionHead) run by the REFERENCE class in the build container."""
    from multimodal_dataset_distillation_b200 import ops
    z = np.load(os.path.join(GOLDEN_DIR, "reparam_small.npz"))
    got = ops.proj_head_forward(torch.from_numpy(z["theta"]).cuda().unsqueeze(0), torch.from_numpy(z["x"]).cuda(), 20)
    assert rel_err(got, z["out"]) < RTOL


@pytest.mark.parametrize("B,dt,d,scale,drop", [(8, 10, 16, 2.0, False), (24, 32, 48, 14.2857, True),
                                               (100, 768, 2304, 2.6593, False), (100, 768, 2304, 14.2857, True),
                                               (100, 768, 2304, 0.1, False)])
def test_contrastive_step_config2(B, dt, d, scale, drop):
    from multimodal_dataset_distillation_b200 import ops
    pr = R.make_problem(N=B, B=B, K=1, dt=dt, d=d, seed=5, dropout=drop, scale=scale)
    mask = pr["masks"][0] if drop else None
    th = pr["theta0"].double().requires_grad_(True)
    Y = pr["Y"].double().requires_grad_(True)
    U = pr["U"].double().requires_grad_(True)
    sc = pr["scale"].double().requires_grad_(True)
    loss = R.infonce(R.row_normalise(U), R.row_normalise(R.head_forward(th, Y, dt, d, None if mask is None else mask.double())), sc)
    gth, gY, gU, gsc = torch.autograd.grad(loss, (th, Y, U, sc))
    got = ops.contrastive_step(pr["theta0"].cuda(), pr["Y"].cuda(), pr["U"].cuda(), pr["scale"].cuda(),
                               None if mask is None else mask.cuda())
    assert abs(float(got["loss"]) - float(loss)) <= RTOL * abs(float(loss))
    assert rel_err(got["g_theta"], gth) < RTOL
    assert rel_err(got["dY"], gY) < RTOL
    assert rel_err(got["dU"], gU) < RTOL
    assert abs(float(got["dscale"]) - float(gsc)) <= RTOL * abs(float(gsc)) + 1e-9


def _run_engine(pr, want_theta_K=True):
    from multimodal_dataset_distillation_b200 import ops
    c = to_cuda(pr)
    return ops.unrolled_match(c["theta0"], c["theta_tgt"], c["Y"], c["U"], c["lr"], c["scale"], c["perms"], c["masks"],
                              want_theta_K=want_theta_K)


@pytest.mark.parametrize("name", ["small_nodrop", "small_drop", "mid_full_batch"])
def test_unrolled_match_small_goldens(golden, name):
    """Goldens made by the reference's ReparamModule + torch double-backward (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN_DIR, "distill_small.npz"))
    g = golden["distill"][f"{name}_f64"]
    pr = R.make_problem(dtype=torch.float32, **g["kw"])
    res = _run_engine(pr)
    out5 = res["out5"].cpu()
    assert abs(out5[2].item() - g["loss"]) <= RTOL * abs(g["loss"])
    assert abs(out5[0].item() - g["num"]) <= RTOL * abs(g["num"])
    assert abs(out5[1].item() - g["den"]) <= RTOL * abs(g["den"])
    assert abs(out5[3].item() - g["dlr"]) <= RTOL * abs(g["dlr"])
    assert abs(out5[4].item() - g["dscale"]) <= RTOL * abs(g["dscale"])
    np.testing.assert_allclose(res["ce"].cpu().numpy(), np.asarray(g["ce"]), rtol=RTOL)
    assert rel_err(res["dY"], z[f"{name}_f64_dY"]) < RTOL
    assert rel_err(res["dU"], z[f"{name}_f64_dU"]) < RTOL
    assert rel_err(res["theta_K"], z[f"{name}_f64_thetaK"]) < RTOL


@pytest.mark.parametrize("key", ["flickr_upstream_f64", "flickr_fork_f64", "flickr_eval_f64", "flickr_drop_f64"])
def test_unrolled_match_flickr_goldens(golden, key):
    """Config 3 (N=B=100, K=8, 768->2304): fp32 engine vs the reference mechanism run in float64."""
    z = np.load(os.path.join(GOLDEN_DIR, "distill_flickr.npz"))
    g = golden["distill"][key]
    pr = R.make_problem(dtype=torch.float32, **g["kw"])
    res = _run_engine(pr)
    out5 = res["out5"].cpu()
    assert abs(out5[2].item() - g["loss"]) <= RTOL * abs(g["loss"])
    assert abs(out5[1].item() - g["den"]) <= RTOL * abs(g["den"])
    assert abs(out5[3].item() - g["dlr"]) <= RTOL * abs(g["dlr"]), (out5[3].item(), g["dlr"])
    assert abs(out5[4].item() - g["dscale"]) <= RTOL * abs(g["dscale"]), (out5[4].item(), g["dscale"])
    np.testing.assert_allclose(res["ce"].cpu().numpy(), np.asarray(g["ce"]), rtol=RTOL)
    rows = slice(None) if key == "flickr_upstream_f64" else slice(None, None, 10)
    assert rel_err(res["dY"][rows], z[key + "_dY"]) < RTOL
    assert rel_err(res["dU"][rows], z[key + "_dU"]) < RTOL
    if "thetaK_sample" in g:
        tk = res["theta_K"].double().cpu()
        assert abs(float(tk.sum()) - g["thetaK_sum"]) <= 1e-4 * abs(g["thetaK_sqsum"]) ** 0.5
        np.testing.assert_allclose(tk[::700001].numpy(), np.asarray(g["thetaK_sample"]), rtol=1e-4, atol=1e-6)


def test_unrolled_match_fp32_reference_behaviour(golden):
    """The reference's own fp32 run (torch CPU) and the fp32 engine both sit within 1e-4 of the f64 truth."""
    g32, g64 = golden["distill"]["flickr_upstream_f32"], golden["distill"]["flickr_upstream_f64"]
    pr = R.make_problem(dtype=torch.float32, **g32["kw"])
    out5 = _run_engine(pr, want_theta_K=False)["out5"].cpu()
    for i, k in ((2, "loss"), (3, "dlr"), (4, "dscale")):
        assert abs(out5[i].item() - g32[k]) <= RTOL * abs(g64[k])


@pytest.mark.parametrize("N,B,K,drop", [(500, 100, 2, False), (120, 100, 3, True), (100, 100, 0, False), (64, 1, 2, False)])
def test_unrolled_match_subsets_vs_oracle(N, B, K, drop):
    """COCO-shaped minibatches (B < N: true random subsets, distill.py:511), K=0 and B=1 edge cases."""
    pr = R.make_problem(N=N, B=B, K=max(K, 1), dt=64, d=96, seed=21, dropout=drop, lr=0.2, scale=5.0, tgt_eps=0.05)
    if K == 0:
        pr["perms"] = pr["perms"][:0]
        pr["masks"] = None
    ref = R.unrolled_match_manual(**{k: (v.double() if isinstance(v, torch.Tensor) and v.is_floating_point() else v)
                                     for k, v in pr.items()})
    res = _run_engine(pr)
    out5 = res["out5"].cpu()
    assert abs(out5[2].item() - float(ref.loss)) <= RTOL * abs(float(ref.loss))
    if K > 0:
        assert abs(out5[3].item() - float(ref.dlr)) <= RTOL * abs(float(ref.dlr)) + 1e-12
        assert abs(out5[4].item() - float(ref.dscale)) <= RTOL * abs(float(ref.dscale)) + 1e-12
        assert rel_err(res["dY"], ref.dY) < RTOL
        assert rel_err(res["dU"], ref.dU) < RTOL
    else:
        assert float(res["dY"].abs().max()) == 0.0 and float(res["dU"].abs().max()) == 0.0
    assert rel_err(res["theta_K"], ref.theta_K) < RTOL


def test_unrolled_match_is_deterministic():
    pr = R.make_problem(N=100, B=100, K=2, dt=768, d=2304, seed=3)
    a, b = _run_engine(pr), _run_engine(pr)
    assert torch.equal(a["out5"], b["out5"]) and torch.equal(a["dY"], b["dY"]) and torch.equal(a["dU"], b["dU"])


def test_autograd_function_and_outer_step():
    """UnrolledMatch.apply + backward() == engine grads; outer SGD(momentum .5) == torch.optim.SGD (distill.py:603-613)."""
    from multimodal_dataset_distillation_b200 import distill
    args = distill.parse_args(["--syn_steps", "2", "--expert_epochs", "1", "--max_start_epoch", "2", "--num_queries", "16",
                               "--mini_batch_size", "16", "--lr_img", "10", "--lr_txt", "10", "--lr_lr", "0.01",
                               "--logit_scale_mode", "fork", "--student_dropout", "0"])
    dt, d = 24, 40
    experts = distill.synthetic_experts(1, 3, dt, d, seed=1, step=0.05).cuda()
    g = torch.Generator().manual_seed(0)
    img, txt = torch.randn(16, d, generator=g), torch.randn(16, dt, generator=g)
    eng = distill.DistillEngine(img, txt, experts, args)
    perms = torch.stack([torch.randperm(16, generator=g) for _ in range(2)])
    # oracle with the fork's aliasing: scale == syn_lr_img
    Y = txt.double().requires_grad_(True); U = img.double().requires_grad_(True)
    lr_i = torch.tensor(0.1, dtype=torch.float64, requires_grad=True); lr_t = torch.tensor(0.1, dtype=torch.float64, requires_grad=True)
    ref = R.unrolled_match_autograd(experts[0, 0].double().cpu(), experts[0, 1].double().cpu(), Y, U, lr_t, lr_i, perms, None, dt, d)
    loss = eng.segment_loss(0, 0, perms)
    assert abs(float(loss) - float(ref.loss)) <= RTOL * float(ref.loss)
    Y0, U0 = eng.Y.detach().clone(), eng.U.detach().clone()
    eng.outer_step(loss)
    assert rel_err(eng.Y.grad, ref.dY) < RTOL and rel_err(eng.U.grad, ref.dU) < RTOL
    assert abs(float(eng.syn_lr_txt.grad) - float(ref.dlr)) <= RTOL * abs(float(ref.dlr))
    assert abs(float(eng.syn_lr_img.grad) - float(ref.dscale)) <= RTOL * abs(float(ref.dscale))
    torch.testing.assert_close(eng.Y.detach(), Y0 - 10 * eng.Y.grad, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(eng.U.detach(), U0 - 10 * eng.U.grad, rtol=1e-5, atol=1e-6)
    assert abs(float(eng.syn_lr_txt) - (0.1 - 0.01 * float(ref.dlr))) < 1e-6


def test_train_mode_dropout_masks_are_fresh_scaled_and_replayed():
    """Students run in train mode (distill.py:446-447): a new dropout-0.1 mask per call, same mask in the reverse sweep."""
    from multimodal_dataset_distillation_b200 import distill, ops
    args = distill.parse_args(["--syn_steps", "2", "--expert_epochs", "1", "--max_start_epoch", "2", "--num_queries", "32",
                               "--mini_batch_size", "32", "--logit_scale_mode", "upstream", "--student_dropout", "0.1"])
    dt, d = 64, 128
    experts = distill.synthetic_experts(1, 3, dt, d, seed=1, step=0.05).cuda()
    g = torch.Generator().manual_seed(0)
    img, txt = torch.randn(32, d, generator=g), torch.randn(32, dt, generator=g)
    eng = distill.DistillEngine(img, txt, experts, args)
    perms = torch.stack([torch.randperm(32, generator=g) for _ in range(2)])
    l1 = eng.segment_loss(0, 0, perms)
    m1 = eng.ws.masks.clone()
    vals = torch.unique(m1).cpu().tolist()
    assert len(vals) == 2 and vals[0] == 0.0 and abs(vals[1] - 1 / 0.9) < 1e-6
    assert 0.05 < float((m1 == 0).float().mean()) < 0.15
    l2 = eng.segment_loss(0, 0, perms)
    assert not torch.equal(m1, eng.ws.masks) and float(l1) != float(l2)             # fresh draw per call
    # the oracle with the SAME masks reproduces loss and gradients (mask replayed in the reverse sweep)
    m2 = eng.ws.masks.clone()
    eng.outer_step(l2)
    ref = R.unrolled_match_autograd(experts[0, 0].double().cpu(), experts[0, 1].double().cpu(), txt.double(), img.double(),
                                    torch.tensor(0.1, dtype=torch.float64), torch.tensor(ops.LOGIT_SCALE_UPSTREAM, dtype=torch.float64),
                                    perms, m2.double().cpu(), dt, d)
    assert abs(float(l2) - float(ref.loss)) <= RTOL * float(ref.loss)
    assert rel_err(eng.Y.grad, ref.dY) < RTOL and rel_err(eng.U.grad, ref.dU) < RTOL


def test_segment_prefetcher_streams_host_experts():
    from multimodal_dataset_distillation_b200 import distill
    host = torch.randn(3, 4, 1000)
    pre = distill.SegmentPrefetcher(host, "cuda")
    pre.prefetch(0, 1, 2)
    for i in range(6):
        sl = pre.get()
        e, s = i % 3, (i + 1) % 2
        assert torch.equal(sl["th0"].cpu(), host[e, s]) and torch.equal(sl["tgt"].cpu(), host[e, s + 2])
        pre.release(sl)
        pre.prefetch((i + 1) % 3, (i + 2) % 2, 2)


@pytest.mark.parametrize("N,B,K,drop", [(500, 100, 2, False), (500, 500, 1, True), (256, 128, 2, False)])
def test_unrolled_match_coco_shape_full_dims(N, B, K, drop):
    """BASELINE.json configs[3]: COCO-shaped distillation (500 pairs; minibatches of 100 and the whole set) at the real
    768 -> 2304 head, so the tensor-core path runs with several M tiles and B x B logits larger than one tile."""
    pr = R.make_problem(N=N, B=B, K=K, dt=768, d=2304, seed=31, dropout=drop, lr=0.1, scale=2.6593, tgt_eps=0.01)
    ref = R.unrolled_match_manual(**{k: (v.double() if isinstance(v, torch.Tensor) and v.is_floating_point() else v)
                                     for k, v in pr.items()})
    res = _run_engine(pr)
    out5 = res["out5"].cpu()
    assert abs(out5[2].item() - float(ref.loss)) <= RTOL * abs(float(ref.loss))
    assert abs(out5[3].item() - float(ref.dlr)) <= RTOL * abs(float(ref.dlr)) + 1e-12
    assert abs(out5[4].item() - float(ref.dscale)) <= RTOL * abs(float(ref.dscale)) + 1e-12
    assert rel_err(res["dY"], ref.dY) < RTOL
    assert rel_err(res["dU"], ref.dU) < RTOL
    assert rel_err(res["theta_K"], ref.theta_K) < RTOL


def test_segments_in_flight_equal_sequential_sum():
    """Throughput mode (DistillEngine.segments_step): two segments on two streams and workspaces give exactly the gradients
    of the two segments run one after the other and summed (a + b is commutative bit for bit), then one outer update."""
    import types
    from multimodal_dataset_distillation_b200 import distill
    N, B, K, dt, d = 48, 32, 3, 64, 96
    args = distill.parse_args(["--syn_steps", str(K), "--expert_epochs", "1", "--max_start_epoch", "2", "--num_queries", str(N),
                               "--mini_batch_size", str(B), "--lr_img", "10", "--lr_txt", "10", "--lr_lr", "0.01",
                               "--logit_scale_mode", "upstream", "--student_dropout", "0.0"])
    g = torch.Generator().manual_seed(4)
    U, Y = torch.randn(N, d, generator=g), torch.randn(N, dt, generator=g)
    experts = distill.synthetic_experts(2, 3, dt, d, seed=1).cuda()
    perms = [torch.stack([torch.randperm(N, generator=g)[:B] for _ in range(K)]) for _ in range(2)]
    segs = [(0, 0), (1, 1)]
    a = distill.DistillEngine(U, Y, experts, args, "cuda")
    la = a.segments_step(segs, perms)
    b = distill.DistillEngine(U, Y, experts, args, "cuda")
    lb = [b.segment_loss(e, s, p) for (e, s), p in zip(segs, perms)]
    b.outer_step(lb[0] + lb[1])
    assert all(torch.equal(x.detach(), y.detach()) for x, y in zip(la, lb))
    assert torch.equal(a.U.detach(), b.U.detach()) and torch.equal(a.Y.detach(), b.Y.detach())
    assert torch.equal(a.syn_lr_txt.detach(), b.syn_lr_txt.detach())


def test_distill_main_from_reference_format_files(tmp_path, capsys):
    """distill.main end to end from files laid out like the reference's: txt_replay_buffer_{n}.pt = list[expert] of
    list[snapshot] of list[param tensors] (buffer.py:64-68, 104-112) and an .npz of frozen image / text embeddings."""
    from multimodal_dataset_distillation_b200 import distill
    dt, d, M = 24, 40, 64
    g = torch.Generator().manual_seed(3)
    shapes = [(d, dt), (d,), (d, d), (d,), (d,), (d,)]                      # ReparamModule order of ProjectionHead
    for n in range(2):
        experts = []
        for _ in range(2):
            snap = [torch.randn(*s, generator=g) * 0.1 for s in shapes]
            traj = [snap]
            for _ in range(3):
                traj.append([p + 0.01 * torch.randn(p.shape, generator=g) for p in traj[-1]])
            experts.append(traj)
        torch.save(experts, tmp_path / f"txt_replay_buffer_{n}.pt")
    np.savez(tmp_path / "embeds.npz", image_embed=torch.randn(M, d, generator=g).numpy(),
             text_embed=torch.randn(M, dt, generator=g).numpy())
    flat = distill.load_expert_buffers(str(tmp_path), "txt")
    assert tuple(flat.shape) == (4, 4, d * dt + d + d * d + 3 * d) and flat.is_cuda
    args = distill.parse_args(["--buffer_path", str(tmp_path), "--embed_path", str(tmp_path / "embeds.npz"), "--num_queries", "16",
                               "--mini_batch_size", "16", "--syn_steps", "3", "--expert_epochs", "1", "--max_start_epoch", "2",
                               "--Iteration", "12", "--lr_img", "1", "--lr_txt", "1", "--lr_lr", "0.001"])
    eng = distill.main(args)
    out = capsys.readouterr().out
    assert "iter = 0000" in out and "iter = 0010" in out
    assert torch.isfinite(eng.Y).all() and torch.isfinite(eng.U).all()
    assert eng.experts.shape[0] == 2 and eng.expert_idx == 13 % 2           # --max_files 1 (distill.py:630): experts of file 0, consumed in order
    # the fork's logit scale is the learnable syn_lr_img: it has received gradient and moved (distill.py:548)
    assert float(eng.syn_lr_img) != float(args.lr_teacher_img)
    with pytest.raises(AssertionError):
        distill.load_expert_buffers(str(tmp_path / "nothing_here"), "txt")


def test_buffer_writes_reference_format_and_distill_reads_it(tmp_path, capsys):
    """buffer.main (text-head experts on frozen embeddings, kernel-backed loss) -> txt_replay_buffer_{n}.pt in the
    reference layout -> distill.load_expert_buffers / distill.main consume it; the teacher's loss goes down."""
    from multimodal_dataset_distillation_b200 import buffer, distill
    dt, d, M = 24, 40, 256
    g = torch.Generator().manual_seed(5)
    img = torch.randn(M, d, generator=g)
    txt = img[:, :dt] * 0.8 + 0.3 * torch.randn(M, dt, generator=g)
    ti = torch.randn(20, d, generator=g)
    tt = ti.repeat_interleave(5, 0)[:, :dt] * 0.8 + 0.3 * torch.randn(100, dt, generator=g)
    np.savez(tmp_path / "embeds.npz", image_embed=img.numpy(), text_embed=txt.numpy(), test_image_embed=ti.numpy(),
             test_text_embed=tt.numpy())
    args = buffer.build_parser().parse_args(["--buffer_path", str(tmp_path / "buffers"), "--embed_path", str(tmp_path / "embeds.npz"),
                                             "--num_experts", "2", "--train_epochs", "4", "--batch_train", "64",
                                             "--lr_teacher_txt", "0.1", "--image_encoder", "nfnet", "--decay"])
    files = buffer.main(args)
    out = capsys.readouterr().out
    assert len(files) == 2 and all(os.path.basename(f) == f"txt_replay_buffer_{i}.pt" for i, f in enumerate(files))
    assert files[0].startswith(os.path.join(str(tmp_path / "buffers"), "flickr", "nfnet", "bert")) and "R@Mean" in out
    traj = torch.load(files[0])
    assert isinstance(traj, list) and len(traj) == 1 and len(traj[0]) == 5            # 1 expert, initial + 4 epoch snapshots
    assert [tuple(p.shape) for p in traj[0][0]] == [(d, dt), (d,), (d, d), (d,), (d,), (d,)] and not traj[0][0][0].is_cuda
    flat = distill.load_expert_buffers(os.path.dirname(files[0]), "txt", max_files=2)
    assert tuple(flat.shape) == (2, 5, d * dt + d + d * d + 3 * d)
    assert not torch.equal(flat[0, 0], flat[0, 4]) and not torch.equal(flat[0, 0], flat[1, 0])
    # the teacher learns: InfoNCE of the last snapshot on the training pairs is below the first one's
    from multimodal_dataset_distillation_b200 import ops
    l0 = float(ops.clip_loss(flat[0, 0], txt[:64].cuda(), img[:64].cuda())["loss"])
    l4 = float(ops.clip_loss(flat[0, 4], txt[:64].cuda(), img[:64].cuda())["loss"])
    assert l4 < l0
    dargs = distill.parse_args(["--buffer_path", os.path.dirname(files[0]), "--embed_path", str(tmp_path / "embeds.npz"),
                                "--num_queries", "32", "--mini_batch_size", "32", "--syn_steps", "2", "--expert_epochs", "1",
                                "--max_start_epoch", "3", "--Iteration", "3", "--max_files", "2", "--lr_img", "1", "--lr_txt", "1",
                                "--eval_it", "2", "--num_eval", "2", "--epoch_eval_train", "3", "--batch_train", "16"])
    eng = distill.main(dargs)
    assert eng.experts.shape[0] == 2 and torch.isfinite(eng.Y).all()
    # evaluation block (distill.py:293-330): iterations 0 and 2, two fresh models each, the nine reference keys
    assert [it for it, _ in eng.eval_history] == [0, 2] and all(len(r) == 2 for _, r in eng.eval_history)
    assert list(eng.eval_history[0][1][0].keys()) == list(ops.RESULT_KEYS)
    assert "Evaluate_01: Img R@1" in capsys.readouterr().out


def test_out_of_range_minibatch_index_poisons_the_loss():
    """The reference raises IndexError on a bad minibatch index (distill.py:512-513).  Here no memory outside the operands is
    touched (the index is clamped) and the call reports NaN losses -- the NaN guard of distill.py:599-600 then stops the run."""
    from multimodal_dataset_distillation_b200 import ops
    pr = R.make_problem(N=16, B=8, K=2, dt=24, d=40, seed=0)
    c = to_cuda(pr)
    good = ops.unrolled_match(c["theta0"], c["theta_tgt"], c["Y"], c["U"], c["lr"], c["scale"], c["perms"], None)
    assert torch.isfinite(good["out5"]).all()
    for bad_value in (16, -1, 10 ** 12):
        perms = c["perms"].clone()
        perms[1, 3] = bad_value
        res = ops.unrolled_match(c["theta0"], c["theta_tgt"], c["Y"], c["U"], c["lr"], c["scale"], perms, None)
        assert torch.isnan(res["out5"]).all()
    again = ops.unrolled_match(c["theta0"], c["theta_tgt"], c["Y"], c["U"], c["lr"], c["scale"], c["perms"], None)
    assert torch.equal(again["out5"], good["out5"]) and torch.equal(again["dY"], good["dY"])     # the flag does not stick
