"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU paths (sharding, candidate merge, packed all-reduce).

The device primitives are replaced by a numpy backend built on the oracle; the collectives and merge rules are the
product code in multimodal_dataset_distillation_b200/dist.py.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import retrieval_ref as RR


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class NumpyBackend:
    """Oracle stand-ins for the three CUDA primitives (tests only)."""

    def scores(self, img, txt_shard, scale):
        S = (np.float32(scale) * img.numpy()) @ txt_shard.numpy().T
        return torch.from_numpy(S.astype(np.float32))

    def best_gt(self, s_i2t, lo, gt_ptr, gt_idx):
        S, ptr, idx = s_i2t.numpy(), gt_ptr.numpy(), gt_idx.numpy()
        bs = np.full(S.shape[0], -np.inf, dtype=np.float32)
        bi = np.full(S.shape[0], -1, dtype=np.int32)
        for r in range(S.shape[0]):
            for c in idx[ptr[r]:ptr[r + 1]]:
                cl = int(c) - lo
                if 0 <= cl < S.shape[1]:
                    s = S[r, cl]
                    if bi[r] < 0 or s > bs[r] or (s == bs[r] and c < bi[r]):
                        bs[r], bi[r] = s, c
        return torch.from_numpy(bs), torch.from_numpy(bi)

    def count(self, s_i2t, lo, thr_s, thr_i):
        S = s_i2t.numpy()
        out = np.zeros(S.shape[0], dtype=np.int32)
        cols = np.arange(S.shape[1]) + lo
        for r in range(S.shape[0]):
            if thr_i[r] >= 0:
                out[r] = np.count_nonzero(S[r] > float(thr_s[r])) + np.count_nonzero((S[r] == float(thr_s[r])) & (cols < int(thr_i[r])))
        return torch.from_numpy(out)

    def ranks_t2i(self, s_i2t, txt2img_shard):
        return torch.from_numpy(RR.ranks_t2i(np.ascontiguousarray(s_i2t.numpy().T), txt2img_shard.numpy()))


def _make_case(seed, n_img, caps, dim, quant):
    img, txt = RR.synthetic_retrieval(n_img, caps, dim, seed=seed)
    if quant:   # heavy ties, including ties between ground-truth captions that live on different shards
        img = np.round(img * 4) / 4
        txt = np.round(txt * 4) / 4
    return img.astype(np.float32), txt.astype(np.float32)


def _worker(rank, world, port, seed, n_img, caps, dim, quant, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multimodal_dataset_distillation_b200 import dist as D
        img, txt = _make_case(seed, n_img, caps, dim, quant)
        T = txt.shape[0]
        txt2img, img2txt = RR.flickr_maps(n_img, caps)
        t2i = np.array([txt2img[t] for t in range(T)], dtype=np.int32)
        ptr = np.arange(0, T + 1, caps, dtype=np.int32)
        idx = np.arange(T, dtype=np.int32)
        lo, hi = D.shard_bounds(T, world, rank)
        r_i, r_t = D.sharded_ranks(torch.from_numpy(img), torch.from_numpy(txt[lo:hi]), lo, torch.from_numpy(t2i[lo:hi]),
                                   torch.from_numpy(ptr), torch.from_numpy(idx), 14.285714, backend=NumpyBackend())
        res = D.sharded_result(r_i, r_t, T)
        # packed all-reduce: every rank contributes rank+1
        a, b, c = torch.full((3, 4), rank + 1.0), torch.full((5,), 10.0 * (rank + 1)), torch.tensor([rank + 0.5, 2.0])
        D.allreduce_packed([a, b, c])
        ret[rank] = dict(r_i=r_i.numpy().copy(), r_t=r_t.numpy().copy(), lo=lo, hi=hi, res=res,
                         packed=(a.clone(), b.clone(), c.clone()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_img,caps,dim,quant", [(23, 5, 16, False), (17, 3, 8, True), (9, 1, 8, True)])
def test_sharded_retrieval_and_packed_allreduce_world2(n_img, caps, dim, quant):
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, 7, n_img, caps, dim, quant, ret), nprocs=world, join=True)
        out = dict(ret)
    img, txt = _make_case(7, n_img, caps, dim, quant)
    S = ((np.float32(14.285714) * img) @ txt.T).astype(np.float32)
    txt2img, img2txt = RR.flickr_maps(n_img, caps)
    ref_i, ref_t = RR.ranks_i2t(S, img2txt), RR.ranks_t2i(np.ascontiguousarray(S.T), txt2img)
    ref = RR.recall_dict(ref_i, ref_t)
    got_t = np.concatenate([out[r]["r_t"] for r in range(world)])
    for r in range(world):
        assert np.array_equal(out[r]["r_i"], ref_i), f"rank {r}"      # identical on every rank, bit-exact
        assert out[r]["res"] == ref
        a, b, c = out[r]["packed"]
        assert torch.equal(a, torch.full((3, 4), 3.0)) and torch.equal(b, torch.full((5,), 30.0))
        assert torch.equal(c, torch.tensor([2.0, 4.0]))
    assert np.array_equal(got_t, ref_t)
    assert out[0]["hi"] == out[1]["lo"] and out[0]["lo"] == 0 and out[1]["hi"] == txt.shape[0]


def test_shard_bounds_cover_exactly():
    from multimodal_dataset_distillation_b200 import dist as D
    for n in (0, 1, 7, 5000, 125000):
        for world in (1, 2, 3, 4, 8):
            spans = [D.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_merge_candidates_tie_rule():
    from multimodal_dataset_distillation_b200 import dist as D
    s = torch.tensor([[1.0, 5.0, -1.0, 2.0], [1.0, 7.0, 0.0, 2.0]])
    i = torch.tensor([[4, 1, -1, 9], [2, 8, -1, 3]], dtype=torch.int32)
    bs, bi = D.merge_candidates(s, i)
    assert bs.tolist()[:2] == [1.0, 7.0] and bi.tolist() == [2, 8, -1, 3]


def test_ranks_sample_disjoint_segments_and_different_minibatches():
    """distill.main under torchrun: rank r starts at expert r and strides by the world size, and draws its own minibatch
    permutations and start epochs; the same (seed, rank) reproduces.  (The collective itself is driven with the real
    engine in tests/test_gpu_dist_engine.py; this is the host-side sampling state only, no device needed.)"""
    import types
    from multimodal_dataset_distillation_b200 import distill

    def sampler(rank, world, seed=3, n_experts=8):
        e = object.__new__(distill.DistillEngine)
        e.args = types.SimpleNamespace(max_start_epoch=4, expert_epochs=1)
        e.experts = torch.empty(n_experts, 6, 1)
        e.N, e.B, e.K = 20, 12, 3
        e._init_sampling(seed, rank, world, n_experts)
        return e

    a, b, a2 = sampler(0, 2), sampler(1, 2), sampler(0, 2)
    sa = [a.sample_segment() for _ in range(6)]
    sb = [b.sample_segment() for _ in range(6)]
    assert [e for e, _ in sa] == [0, 2, 4, 6, 0, 2] and [e for e, _ in sb] == [1, 3, 5, 7, 1, 3]
    assert all(0 <= s < 4 for _, s in sa + sb) and [s for _, s in sa] != [s for _, s in sb]
    pa, pb = a.draw_perms(), b.draw_perms()
    assert tuple(pa.shape) == (3, 12) and not torch.equal(pa, pb)
    assert all(len(set(row.tolist())) == 12 for row in pa)                  # randperm(N)[:B]: unique indices per step
    assert [a2.sample_segment() for _ in range(6)] == sa and torch.equal(a2.draw_perms(), pa)
    one = sampler(0, 1)
    assert [one.sample_segment()[0] for _ in range(4)] == [0, 1, 2, 3]      # world 1: the reference's in-order walk
