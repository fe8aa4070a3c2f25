"""CPU: host-side mirrors of the reference interface (CLI, ReparamModule, map conversion, buffer format)."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import GOLDEN_DIR


def test_cli_matches_reference_flags():
    """Every flag of distill.py:625-679 exists with the same default (tests/golden/cli.json, extracted with ast)."""
    from multimodal_dataset_distillation_b200 import distill
    with open(os.path.join(GOLDEN_DIR, "cli.json")) as f:
        ref = json.load(f)["flags"]
    parser = distill.build_parser()
    actions = {a.option_strings[0]: a for a in parser._actions if a.option_strings}
    assert len(ref) == 55
    assert "--segments_in_flight" in actions and actions["--segments_in_flight"].default == 1     # extension, off by default
    for e in ref:
        a = actions.get(e["flag"])
        assert a is not None, e["flag"]
        if e.get("action") == "store_true":
            assert a.default is False and a.nargs == 0
        if "default" in e and e["default"] != "<non-literal>":
            assert a.default == e["default"], e["flag"]
        if "type" in e and e["type"] in ("int", "float", "str", "bool"):
            assert a.type is {"int": int, "float": float, "str": str, "bool": bool}[e["type"]], e["flag"]
        if "choices" in e:
            assert set(e["choices"]) <= set(a.choices), e["flag"]
    args = distill.parse_args(["--syn_steps", "8", "--bogus_flag", "1"])        # parse_known_args leniency, distill.py:680
    assert args.syn_steps == 8 and args.lr_teacher_img == 0.1


class Head(nn.Module):
    def __init__(self, e, p):
        super().__init__()
        self.projection = nn.Linear(e, p)
        self.gelu = nn.GELU()
        self.fc = nn.Linear(p, p)
        self.dropout = nn.Dropout(0.1)
        self.layer_norm = nn.LayerNorm(p)

    def forward(self, x):
        pr = self.projection(x)
        return self.layer_norm(self.dropout(self.fc(self.gelu(pr))) + pr)


def test_reparam_module_interface(golden):
    from multimodal_dataset_distillation_b200.reparam_module import ReparamModule
    rp = ReparamModule(Head(768, 2304))
    g = golden["reparam"]
    assert rp.param_numel == g["param_numel"]
    assert [f"{mn}.{n}" for mn, n in rp._param_infos] == g["names"]
    assert list(rp._param_numels) == g["numels"] and [list(s) for s in rp._param_shapes] == g["shapes"]
    assert isinstance(rp.flat_param, nn.Parameter) and [n for n, _ in rp.named_parameters()] == ["flat_param"]
    z = np.load(os.path.join(GOLDEN_DIR, "reparam_small.npz"))
    small = ReparamModule(Head(12, 20)).eval()
    x, th = torch.from_numpy(z["x"]), torch.from_numpy(z["theta"])
    for fp in (th, th.unsqueeze(0)):                                   # [P] and the DataParallel [1,P] slice
        np.testing.assert_allclose(small(x, flat_param=fp).detach().numpy(), z["out"], rtol=1e-6, atol=1e-6)
    # differentiable w.r.t. the external flat vector, twice (create_graph=True as in distill.py:565)
    th = th.clone().requires_grad_(True)
    out = small(x, flat_param=th).pow(2).sum()
    (g1,) = torch.autograd.grad(out, th, create_graph=True)
    g1.sum().backward()
    assert th.grad is not None and th.grad.abs().sum() > 0
    # views are restored after the call
    assert small.module.projection.weight.data_ptr() == small.flat_param.data_ptr()


def test_reparam_shared_parameters_and_buffers():
    from multimodal_dataset_distillation_b200.reparam_module import ReparamModule

    class Tied(nn.Module):
        def __init__(self):
            super().__init__()
            self.a = nn.Linear(4, 4, bias=False)
            self.b = nn.Linear(4, 4, bias=False)
            self.b.weight = self.a.weight
            self.bn = nn.BatchNorm1d(4)

        def forward(self, x):
            return self.bn(self.b(self.a(x)))
    rp = ReparamModule(Tied())
    assert rp.param_numel == 16 + 8 and len(rp._shared_param_infos) == 1 and len(rp._buffer_infos) == 3
    x = torch.randn(5, 4)
    out = rp(x, flat_param=torch.ones(24), buffers=[torch.zeros(4), torch.ones(4), torch.tensor(0)])
    assert out.shape == (5, 4)


def test_maps_to_arrays():
    from multimodal_dataset_distillation_b200 import ops
    img2txt = {0: [0, 1, 2], 1: [3], 2: [4, 5]}
    txt2img = {0: 0, 1: 0, 2: 0, 3: 1, 4: 2, 5: 2}
    t2i, ptr, idx = ops.maps_to_arrays(txt2img, img2txt, 3, 6)
    assert t2i.tolist() == [0, 0, 0, 1, 2, 2] and ptr.tolist() == [0, 3, 4, 6] and idx.tolist() == [0, 1, 2, 3, 4, 5]
    assert t2i.dtype == ptr.dtype == idx.dtype == np.int32
    # the real Flickr30k test annotation shape: 1000 images x 5 captions in contiguous blocks
    from oracle import retrieval_ref as RR
    t2i, ptr, idx = ops.maps_to_arrays(*RR.flickr_maps(1000, 5), 1000, 5000)
    assert ptr[-1] == 5000 and (np.diff(ptr) == 5).all() and (t2i == np.arange(5000) // 5).all()


def test_expert_buffer_format_roundtrip(tmp_path):
    """buffer.py:104-112 writes list[expert] of list[snapshot] of list[param tensors]; we flatten once."""
    from multimodal_dataset_distillation_b200 import distill, ops
    dt, d = 6, 10
    shapes = [(d, dt), (d,), (d, d), (d,), (d,), (d,)]
    traj = [[[torch.randn(s) for s in shapes] for _snap in range(3)] for _exp in range(2)]
    torch.save(traj, tmp_path / "txt_replay_buffer_0.pt")
    torch.save(traj[:1], tmp_path / "txt_replay_buffer_1.pt")
    flat = distill.load_expert_buffers(str(tmp_path), "txt", None, device="cpu")
    assert flat.shape == (3, 3, ops.head_numel(dt, d))
    assert torch.equal(flat[1, 2], torch.cat([p.reshape(-1) for p in traj[1][2]]))
    with pytest.raises(AssertionError, match="No buffers detected"):
        distill.load_expert_buffers(str(tmp_path / "nope"), "txt", None, device="cpu")


def test_networks_mirror_layout_and_errors(golden):
    """networks.ProjectionHead has the reference's module tree (flat layout == ReparamModule's, golden 'reparam' names);
    CLIPModel_full refuses to guess an image encoder and every kernel-backed call refuses CPU tensors."""
    import types
    from multimodal_dataset_distillation_b200 import networks, ops, infonce, distill
    from multimodal_dataset_distillation_b200.reparam_module import ReparamModule
    head = networks.ProjectionHead(768, 2304)
    rp = ReparamModule(networks.ProjectionHead(768, 2304))
    assert [f"{mn}.{n}" for mn, n in rp._param_infos] == golden["reparam"]["names"]
    flat = head.flat_parameters()
    assert flat.numel() == golden["reparam"]["param_numel"] == ops.head_numel(768, 2304)
    off = 0
    for p in head.parameters():                                   # parameters() order == flat order
        assert torch.equal(flat[off:off + p.numel()], p.detach().reshape(-1))
        off += p.numel()
    head.eval()
    assert head.dropout_mask(4, "cpu") is None
    head.train()
    m = head.dropout_mask(64, "cpu")
    assert m.shape == (64, 2304) and set(torch.unique(m).tolist()) <= {0.0, 1.0 / 0.9} or torch.allclose(
        torch.unique(m), torch.tensor([0.0, 1.0 / 0.9]))
    with pytest.raises(ValueError):
        networks.CLIPModel_full(types.SimpleNamespace(distill=True))
    net = networks.CLIPModel_full(types.SimpleNamespace(distill=True), image_encoder=nn.Identity(), image_embedding=16,
                                  text_embedding=8)
    assert abs(net.logit_scale - 1 / 0.07) < 1e-9
    with pytest.raises(RuntimeError):
        net(torch.randn(4, 16), torch.randn(4, 8), 0)            # CPU tensors: no CPU path
    with pytest.raises(TypeError):
        net(torch.randn(4, 16), ["a caption"] * 4, 0)             # raw captions without a text encoder
    with pytest.raises(RuntimeError):
        infonce.infonce_loss(torch.randn(4, 8), torch.randn(4, 8), 1.0)
    with pytest.raises(RuntimeError):
        ops.nearest_rows(torch.randn(2, 8), torch.randn(5, 8))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            distill.nearest_neighbor(["a", "b"], np.zeros((1, 4), np.float32), np.zeros((2, 4), np.float32))


def test_buffer_cli_matches_reference_flags():
    """Every flag of buffer.py:119-160 exists with the same type and default (tests/golden/cli_buffer.json, extracted with ast)."""
    from multimodal_dataset_distillation_b200 import buffer
    with open(os.path.join(GOLDEN_DIR, "cli_buffer.json")) as f:
        ref = json.load(f)["flags"]
    actions = {a.option_strings[0]: a for a in buffer.build_parser()._actions if a.option_strings}
    assert len(ref) == 35
    for e in ref:
        a = actions.get(e["flag"])
        assert a is not None, e["flag"]
        if e.get("action") == "store_true":
            assert a.default is False and a.nargs == 0
        if "default" in e and e["default"] != "<non-literal>":
            assert a.default == e["default"], e["flag"]
        if "type" in e and e["type"] in ("int", "float", "str", "bool"):
            assert a.type is {"int": int, "float": float, "str": str, "bool": bool}[e["type"]], e["flag"]
        if "choices" in e:
            assert set(e["choices"]) <= set(a.choices), e["flag"]
    args = buffer.build_parser().parse_args(["--dataset", "coco", "--image_encoder", "nfnet"])
    assert buffer.save_dir_of(args) == os.path.join("./buffers", "coco", "nfnet", "bert")       # buffer.py:27-31


def test_maps_to_arrays_rejects_out_of_range_ground_truth():
    from multimodal_dataset_distillation_b200 import ops
    t2i, ptr, idx = ops.maps_to_arrays({0: 0, 1: 0, 2: 1}, {0: [0, 1], 1: [2]}, 2, 3)
    assert t2i.tolist() == [0, 0, 1] and ptr.tolist() == [0, 2, 3] and idx.tolist() == [0, 1, 2]
    with pytest.raises(IndexError):
        ops.maps_to_arrays({0: 0, 1: 2, 2: 1}, {0: [0, 1], 1: [2]}, 2, 3)          # image 2 of 2
    with pytest.raises(IndexError):
        ops.maps_to_arrays({0: 0, 1: 0, 2: 1}, {0: [0, 3], 1: [2]}, 2, 3)          # caption 3 of 3
