"""Round-2 engine pieces through the C ABI: one-pass matching loss + adjoint, the fused outer update, the engine's own
Philox dropout masks (drawn inside the launch graph, replayed in the reverse sweep) and the autograd-free fast path.

Tolerance as in test_gpu_distill.py: 1e-4 relative (norm-wise for tensors) unless the arithmetic is identical, where
equality is asserted bit for bit.
"""
import ctypes as C

import pytest
import torch

from oracle import distill_ref as R

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def rel_err(got, ref):
    got, ref = torch.as_tensor(got).double().cpu(), torch.as_tensor(ref).double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-300))


def rms_rel_err(got, ref):
    got, ref = torch.as_tensor(got).double().cpu(), torch.as_tensor(ref).double().cpu()
    return float((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt().clamp_min(1e-300))


@pytest.mark.parametrize("n", [1, 5, 1023, 7_087_104])
def test_match_final_equals_fwd_plus_bwd(n):
    """distill.py:588-598 + 606: {num, den, num/den} and a = 2 (theta_K - theta*) / den from ONE pass."""
    from multimodal_dataset_distillation_b200 import ops
    from multimodal_dataset_distillation_b200._lib import lib, check
    g = torch.Generator().manual_seed(n)
    thK, tgt, th0 = (torch.randn(n, generator=g).cuda() for _ in range(3))
    ref3 = ops.match_loss(thK, tgt, th0)
    ref_a = ops.match_loss_bwd(thK, tgt, ref3)
    out3, adj = torch.empty(3, device="cuda"), torch.empty(n, device="cuda")
    scratch = torch.zeros(lib().vldd_match_loss_scratch_bytes(), dtype=torch.uint8, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):                                    # twice: the ticket re-arms itself
        check(lib().vldd_match_final(p(thK), p(tgt), p(ref3[1:2]), n, p(out3), p(adj), p(scratch), st), "match_final")
        d = (thK.double() - tgt.double())
        assert abs(float(out3[0]) - float(d.pow(2).sum())) <= 1e-5 * float(d.pow(2).sum())
        assert float(out3[1]) == float(ref3[1])
        assert abs(float(out3[2]) - float(ref3[2])) <= 1e-6 * abs(float(ref3[2]))
        assert rel_err(adj, ref_a) < 1e-6


def test_outer_update_matches_torch_sgd_and_skips_on_nan():
    """distill.py:233-241, 603-613: three torch.optim.SGD(momentum=0.5) steps == one vldd_outer_update launch."""
    from multimodal_dataset_distillation_b200 import ops
    g = torch.Generator().manual_seed(0)
    U, Y = torch.randn(37, 50, generator=g).cuda(), torch.randn(37, 21, generator=g).cuda()
    lr_i, lr_t = torch.tensor(0.1).cuda(), torch.tensor(0.2).cuda()
    pU, pY, pi, pt = (torch.nn.Parameter(t.clone()) for t in (U, Y, lr_i, lr_t))
    opts = [torch.optim.SGD([pU], lr=10.0, momentum=0.5), torch.optim.SGD([pY], lr=7.0, momentum=0.5),
            torch.optim.SGD([pi, pt], lr=0.01, momentum=0.5)]
    bU, bY, bl = torch.zeros_like(U), torch.zeros_like(Y), torch.zeros(2, device="cuda")
    loss, skipped = torch.ones(1, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")
    for it in range(3):
        gU, gY = torch.randn(37, 50, generator=g).cuda(), torch.randn(37, 21, generator=g).cuda()
        gl = torch.randn(2, generator=g).cuda()
        pU.grad, pY.grad, pi.grad, pt.grad = 0.5 * gU, 0.5 * gY, 0.5 * gl[0], 0.5 * gl[1]
        for o in opts:
            o.step()
        ops.outer_update(U, gU, bU, 10.0, Y, gY, bY, 7.0, lr_i, lr_t, gl[0:1], gl[1:2], bl, 0.01, 0.5, it == 0, 0.5, loss, skipped)
        torch.testing.assert_close(U, pU.detach(), rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(Y, pY.detach(), rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(torch.stack([lr_i, lr_t]), torch.stack([pi.detach(), pt.detach()]), rtol=1e-6, atol=1e-7)
    assert int(skipped) == 0
    before = [t.clone() for t in (U, Y, lr_i, lr_t, bU)]
    loss.fill_(float("nan"))
    ops.outer_update(U, gU, bU, 10.0, Y, gY, bY, 7.0, lr_i, lr_t, None, gl[1:2], bl, 0.01, 0.5, False, 1.0, loss, skipped)
    assert int(skipped) == 1 and all(torch.equal(a, b) for a, b in zip(before, (U, Y, lr_i, lr_t, bU)))


def test_philox_masks_reproducible_scaled_and_advancing():
    """networks.py:629,636 nn.Dropout(0.1): values in {0, 1/0.9}, ~10 % zeros, same (seed, draw) -> same mask, next draw differs."""
    from multimodal_dataset_distillation_b200 import ops
    st = ops.make_rng_state(1234, "cuda")
    m1 = ops.dropout_masks((8, 100, 2304), 0.1, st, advance=False)
    m1b = ops.dropout_masks((8, 100, 2304), 0.1, st, advance=True)
    assert torch.equal(m1, m1b) and st.cpu().tolist() == [1234, 1]
    m2 = ops.dropout_masks((8, 100, 2304), 0.1, st)
    assert not torch.equal(m1, m2) and st.cpu().tolist() == [1234, 2]
    vals = torch.unique(m1).cpu().tolist()
    assert len(vals) == 2 and vals[0] == 0.0 and abs(vals[1] - 1 / 0.9) < 1e-6
    frac = float((m1 == 0).float().mean())
    assert abs(frac - 0.1) < 2e-3, frac                    # 1.8 M draws: 3 sigma = 7e-4
    assert abs(float(m1.mean()) - 1.0) < 3e-3
    # no structure along the row: every column's drop rate is near p as well
    col = (m1 == 0).float().mean(dim=(0, 1))
    assert float(col.max()) < 0.16 and float(col.min()) > 0.05
    other_seed = ops.dropout_masks((8, 100, 2304), 0.1, ops.make_rng_state(1235, "cuda"))
    assert not torch.equal(other_seed, m1)
    odd = ops.dropout_masks((3, 7, 5), 0.25, ops.make_rng_state(1, "cuda"))          # n % 4 != 0: tail path
    assert all(v == 0.0 or abs(v - 1 / 0.75) < 1e-6 for v in torch.unique(odd).cpu().tolist()) and odd.numel() == 105


@pytest.mark.parametrize("N,B,K,dt,d", [(32, 32, 2, 64, 128), (100, 100, 3, 768, 2304)])
def test_engine_draws_its_own_masks_and_replays_them(N, B, K, dt, d):
    """Train-mode students (distill.py:446-447): the engine draws the K masks inside its launch graph; the oracle fed with
    the masks read back from the workspace reproduces loss and gradients, so the reverse sweep used the same masks."""
    from multimodal_dataset_distillation_b200 import ops
    pr = R.make_problem(N=N, B=B, K=K, dt=dt, d=d, seed=17, lr=0.1, scale=2.6593)
    c = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in pr.items()}
    st = ops.make_rng_state(99, "cuda")
    ws = ops.UnrollWorkspace(N, B, K, dt, d, "cuda")
    seen = []
    for call in range(3):                                 # call 0 captures the graph, 1 and 2 replay it
        res = ops.unrolled_match(c["theta0"], c["theta_tgt"], c["Y"], c["U"], c["lr"], c["scale"], c["perms"], None, ws,
                                 dropout_p=0.1, rng_state=st)
        masks = ws.masks.clone()
        assert all(not torch.equal(masks, m) for m in seen)          # a fresh draw per call, also on graph replays
        seen.append(masks)
        assert st.cpu().tolist() == [99, call + 1]
        assert torch.equal(masks, ops.dropout_masks((K, B, d), 0.1, torch.tensor([99, call], device="cuda"), advance=False))
        ref = R.unrolled_match_manual(**{k: (v.double() if isinstance(v, torch.Tensor) and v.is_floating_point() else v)
                                         for k, v in dict(pr, masks=masks.cpu()).items()})
        assert abs(float(res["out5"][2]) - float(ref.loss)) <= RTOL * float(ref.loss)
        assert rel_err(res["dY"], ref.dY) < RTOL and rel_err(res["dU"], ref.dU) < RTOL
        assert rms_rel_err(res["dY"], ref.dY) < RTOL and rms_rel_err(res["dU"], ref.dU) < RTOL


def test_segment_and_index_addresses_may_change_between_replays():
    """theta_0, theta* and the minibatch indices reach the replayed launch graph through a device-side pointer table: calls on
    one workspace with DIFFERENT source tensors (other addresses, other contents) must each equal a fresh-workspace call."""
    from multimodal_dataset_distillation_b200 import ops
    N, B, K, dt, d = 40, 24, 3, 64, 96
    prs = [R.make_problem(N=N, B=B, K=K, dt=dt, d=d, seed=s, lr=0.1, scale=2.6593) for s in (5, 6, 7)]
    cs = [{k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in pr.items()} for pr in prs]
    Y, U = cs[0]["Y"], cs[0]["U"]                         # the synthetic set keeps its address (it is part of the graph key)
    ws = ops.UnrollWorkspace(N, B, K, dt, d, "cuda")
    outs = []
    for c in cs + cs[:1]:                                  # call 0 captures, the others replay with new sources
        res = ops.unrolled_match(c["theta0"], c["theta_tgt"], Y, U, c["lr"], c["scale"], c["perms"], None, ws)
        fresh = ops.unrolled_match(c["theta0"], c["theta_tgt"], Y, U, c["lr"], c["scale"], c["perms"], None,
                                   ops.UnrollWorkspace(N, B, K, dt, d, "cuda"))
        for k in ("out5", "dY", "dU"):
            assert torch.equal(res[k], fresh[k]), k
        outs.append(res["dY"].clone())
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])
    assert torch.equal(outs[0], outs[3])
    ref = R.unrolled_match_manual(**{k: (v.double() if isinstance(v, torch.Tensor) and v.is_floating_point() else v)
                                     for k, v in dict(prs[2], Y=prs[0]["Y"], U=prs[0]["U"]).items()})
    assert rel_err(outs[2], ref.dY) < RTOL


@pytest.mark.parametrize("mode", ["fork", "upstream"])
def test_step_fast_equals_autograd_path(mode):
    """DistillEngine.step_fast (engine call + fused update, no torch kernels) == segment_loss + backward + outer_step."""
    from multimodal_dataset_distillation_b200 import distill
    N, B, K, dt, d = 48, 32, 3, 64, 96
    args = distill.parse_args(["--syn_steps", str(K), "--expert_epochs", "1", "--max_start_epoch", "2", "--num_queries", str(N),
                               "--mini_batch_size", str(B), "--lr_img", "10", "--lr_txt", "10", "--lr_lr", "0.01",
                               "--logit_scale_mode", mode, "--student_dropout", "0.0"])
    g = torch.Generator().manual_seed(4)
    U, Y = torch.randn(N, d, generator=g), torch.randn(N, dt, generator=g)
    experts = distill.synthetic_experts(2, 3, dt, d, seed=1).cuda()
    perms = [torch.stack([torch.randperm(N, generator=g)[:B] for _ in range(K)]).cuda() for _ in range(3)]
    a, b = distill.DistillEngine(U, Y, experts, args, "cuda"), distill.DistillEngine(U, Y, experts, args, "cuda")
    for i in range(3):                                    # three iterations: first step, momentum, graph replay
        la = float(a.step_fast(i % 2, i % 2, perms[i]))
        lb = b.segment_loss(i % 2, i % 2, perms[i])
        b.outer_step(lb)
        assert la == float(lb)
        torch.testing.assert_close(a.U.detach(), b.U.detach(), rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(a.Y.detach(), b.Y.detach(), rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(a.syn_lr_txt.detach(), b.syn_lr_txt.detach(), rtol=1e-6, atol=1e-9)
        torch.testing.assert_close(a.syn_lr_img.detach(), b.syn_lr_img.detach(), rtol=1e-6, atol=1e-9)
    assert int(a.ws.skipped) == 0


def test_step_io_feeds_indices_and_returns_every_loss():
    """distill.StepIO: minibatch indices uploaded on a copy stream into rotating device slots, losses fetched through the
    device ring one step late -- the values must be the ones a synchronous loop sees."""
    from multimodal_dataset_distillation_b200 import distill
    N, B, K, dt, d = 48, 32, 2, 64, 96
    args = distill.parse_args(["--syn_steps", str(K), "--expert_epochs", "1", "--max_start_epoch", "2", "--num_queries", str(N),
                               "--mini_batch_size", str(B), "--lr_img", "10", "--lr_txt", "10", "--lr_lr", "0.01",
                               "--student_dropout", "0.0"])
    g = torch.Generator().manual_seed(9)
    U, Y = torch.randn(N, d, generator=g), torch.randn(N, dt, generator=g)
    experts = distill.synthetic_experts(2, 3, dt, d, seed=2).cuda()
    perms_host = [torch.stack([torch.randperm(N, generator=g)[:B] for _ in range(K)]).pin_memory() for _ in range(7)]
    a, b = distill.DistillEngine(U, Y, experts, args, "cuda"), distill.DistillEngine(U, Y, experts, args, "cuda")
    want = [float(b.step_fast(i % 2, i % 2, perms_host[i].cuda())) for i in range(7)]        # synchronous reference loop
    io = distill.StepIO(K, B, "cuda", torch.cuda.Stream())
    io.upload_perms(0, perms_host[0])
    got = []
    for i in range(7):
        p = io.perms_for(i)
        loss = a.step_fast(i % 2, i % 2, p)
        io.step_done(i, loss)
        if i + 1 < 7:
            io.upload_perms(i + 1, perms_host[i + 1])
        if i > 0:
            got.append(io.loss(i - 1))                    # one step late, while step i is in flight
    got.append(io.loss(6))
    assert got == want
    assert torch.equal(a.U.detach(), b.U.detach()) and torch.equal(a.Y.detach(), b.Y.detach())


def test_step_fast_refuses_to_step_on_nan_loss():
    """distill.py:599-600: a NaN loss must not reach the optimiser."""
    from multimodal_dataset_distillation_b200 import distill
    N, B, K, dt, d = 16, 16, 1, 24, 40
    args = distill.parse_args(["--syn_steps", str(K), "--expert_epochs", "1", "--max_start_epoch", "2", "--num_queries", str(N),
                               "--mini_batch_size", str(B), "--logit_scale_mode", "upstream", "--student_dropout", "0.0"])
    g = torch.Generator().manual_seed(4)
    U, Y = torch.randn(N, d, generator=g), torch.randn(N, dt, generator=g)
    experts = distill.synthetic_experts(1, 3, dt, d, seed=1).cuda()
    experts[0, 1, 5] = float("nan")                       # poisoned target snapshot
    eng = distill.DistillEngine(U, Y, experts, args, "cuda")
    U0, Y0 = eng.U.detach().clone(), eng.Y.detach().clone()
    loss = eng.step_fast(0, 0)
    assert not torch.isfinite(loss).item() and int(eng.ws.skipped) == 1
    assert torch.equal(eng.U.detach(), U0) and torch.equal(eng.Y.detach(), Y0)


def test_segment_cache_lru_uploads_only_misses():
    """distill.SegmentCache: device-side LRU of uploaded expert snapshots (the reference re-uploads theta_start / theta_target
    every iteration, distill.py:466-476): hits copy nothing, eviction is least-recently-used, contents are always right."""
    from multimodal_dataset_distillation_b200 import distill
    host = torch.randn(3, 5, 1000)                       # 3 experts x 5 snapshots
    cache = distill.SegmentCache(host, "cuda", capacity=4)
    P4 = 1000 * 4

    def fetch(e, s, k=1):
        cache.prefetch(e, s, k)
        sl = cache.get()
        torch.cuda.current_stream().synchronize()
        assert torch.equal(sl["th0"].cpu(), host[e, s]) and torch.equal(sl["tgt"].cpu(), host[e, s + k])
        cache.release(sl)

    fetch(0, 0)                                          # misses: (0,0), (0,1)
    assert cache.h2d_bytes == 2 * P4 and (cache.hits, cache.misses) == (0, 2)
    fetch(0, 1)                                          # (0,1) hit, (0,2) miss
    assert cache.h2d_bytes == 3 * P4 and (cache.hits, cache.misses) == (1, 3)
    fetch(0, 0)                                          # both resident
    assert cache.h2d_bytes == 3 * P4 and cache.hits == 3
    fetch(1, 0)                                          # 2 misses: 5 snapshots wanted, capacity 4 -> (0,2), the LRU, is evicted
    assert len(cache.slots) == 4 and (0, 2) not in cache.slots and (0, 0) in cache.slots
    fetch(0, 1)                                          # (0,1) still there, (0,2) comes back
    assert torch.equal(cache.slots[(0, 2)]["buf"].cpu(), host[0, 2])
    for i in range(12):                                  # churn through everything: contents stay right (asserted in fetch)
        fetch(i % 3, (i * 2) % 4)


def test_distill_main_writes_the_distilled_set(tmp_path, capsys):
    """distill.main ends with {save_path}/distilled_{it}.pt holding U, Y and the learned student learning rates."""
    from multimodal_dataset_distillation_b200 import distill
    args = distill.parse_args(["--synthetic", "--num_queries", "8", "--mini_batch_size", "8", "--syn_steps", "1", "--expert_epochs", "1",
                               "--max_start_epoch", "2", "--Iteration", "3", "--lr_img", "1", "--lr_txt", "1", "--save_path",
                               str(tmp_path / "out"), "--student_dropout", "0.1"])
    eng = distill.main(args)
    blob = torch.load(eng.saved_to)
    assert blob["iteration"] == 3 and tuple(blob["U"].shape) == (8, 2304) and tuple(blob["Y"].shape) == (8, 768)
    assert torch.equal(blob["U"], eng.U.detach().cpu()) and blob["syn_lr_txt"] == float(eng.syn_lr_txt.detach())
    assert "iter = 0000" in capsys.readouterr().out
