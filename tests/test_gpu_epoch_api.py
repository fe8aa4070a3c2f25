"""The reference-facing Python API of epoch.py on the GPU: epoch_test / epoch_test_metrics / evaluate_synset / itm_eval."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from conftest import GOLDEN_DIR
from oracle import retrieval_ref as RR

pytestmark = pytest.mark.gpu


class Head(nn.Module):          # same module tree as networks.py:625-646
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


class FakeCLIP(nn.Module):
    """Stands in for CLIPModel_full: `.image_encoder`, `.text_projection`, forward(image, caption, epoch) -> (loss, acc)."""

    def __init__(self, dt, d, din):
        super().__init__()
        self.image_encoder = nn.Linear(din, d)
        self.text_projection = Head(dt, d)

    def forward(self, image, caption, epoch):
        x = self.image_encoder(image)
        y = self.text_projection(caption.float())
        x = x / x.norm(dim=1, keepdim=True)
        y = y / y.norm(dim=1, keepdim=True)
        logits = 14.285714 * x @ y.t()
        gt = torch.arange(len(logits), device=logits.device)
        loss = (F.cross_entropy(logits, gt) + F.cross_entropy(logits.t(), gt)) / 2
        acc = ((logits.argmax(1) == gt).sum().item() + (logits.argmax(0) == gt).sum().item()) / 2
        return loss, acc


class FakeSet:
    def __init__(self, n_img, caps):
        self.txt2img, self.img2txt = RR.flickr_maps(n_img, caps)


class FakeLoader(list):
    pass


def _loader(feats, n_img, caps, bs=16):
    ld = FakeLoader((feats[i:i + bs], torch.arange(i, min(i + bs, n_img))) for i in range(0, n_img, bs))
    ld.dataset = FakeSet(n_img, caps)
    return ld


def test_epoch_test_matches_reference_epoch_test_golden():
    """epoch.epoch_test(dataloader, model, device, bert_test_embed) vs epoch_original.epoch_test run on the same model."""
    from multimodal_dataset_distillation_b200 import epoch
    z = np.load(os.path.join(GOLDEN_DIR, "epoch_test_small.npz"))
    I, C, D, dt = (int(x) for x in z["dims"])
    model = FakeCLIP(dt, D, D).cuda()
    model.image_encoder = nn.Identity()
    theta = torch.from_numpy(z["theta"])
    offs = np.cumsum([0, D * dt, D, D * D, D, D, D])
    with torch.no_grad():
        for p, a, b in zip(model.text_projection.parameters(), offs[:-1], offs[1:]):
            p.copy_(theta[a:b].view_as(p))
    loader = _loader(torch.from_numpy(z["feats"]), I, C)
    s1, s2 = epoch.epoch_test(loader, model, "cuda", torch.from_numpy(z["bert"]))
    assert isinstance(s1, np.ndarray) and s1.shape == (I, I * C) and s2.shape == (I * C, I)
    for got, ref in ((s1, z["s_i2t"]), (s2, z["s_t2i"])):
        kg, kr = got > -100, ref > -100
        assert (kg.sum(axis=1) == 128).all() and (kg != kr).sum() <= 4
        np.testing.assert_allclose(got[kg & kr], ref[kg & kr], rtol=2e-5, atol=2e-5)
    # reference's own pipeline: epoch_test -> itm_eval; ours: same matrices and the streaming variant agree on recall
    res_a = epoch.itm_eval(s1, s2, loader.dataset.txt2img, loader.dataset.img2txt)
    res_ref = RR.itm_eval_ref(z["s_i2t"], z["s_t2i"], loader.dataset.txt2img, loader.dataset.img2txt)
    res_b = epoch.epoch_test_metrics(loader, model, "cuda", torch.from_numpy(z["bert"]))
    for k in RR.RESULT_KEYS:
        assert abs(res_a[k] - res_ref[k]) <= 100.0 / I + 1e-9, k
        assert abs(res_b[k] - res_ref[k]) <= 100.0 / I + 1e-9, k


def test_itm_eval_accepts_numpy_cpu_tensors_and_cuda_tensors():
    from multimodal_dataset_distillation_b200 import epoch
    img, txt = RR.synthetic_retrieval(50, 5, 32, seed=3)
    S = ((np.float32(14.285714) * img) @ txt.T).astype(np.float32)
    St = np.ascontiguousarray(S.T)
    txt2img, img2txt = RR.flickr_maps(50, 5)
    want = RR.itm_eval_ref(S, St, txt2img, img2txt)
    assert epoch.itm_eval(S, St, txt2img, img2txt) == want
    assert epoch.itm_eval(torch.from_numpy(S), torch.from_numpy(St), txt2img, img2txt) == want
    assert epoch.itm_eval(torch.from_numpy(S).cuda(), torch.from_numpy(St).cuda(), txt2img, img2txt) == want
    assert list(want.keys()) == list(RR.RESULT_KEYS)


def test_evaluate_synset_signature_and_result():
    """evaluate_synset(it_eval, net, images_train, labels_train, testloader, args, bert_test_embed) -> (net, accs, dict)."""
    import types
    from multimodal_dataset_distillation_b200 import epoch
    torch.manual_seed(0)
    dt, d, din, n_img, caps = 16, 32, 24, 40, 5
    net = FakeCLIP(dt, d, din)
    args = types.SimpleNamespace(device="cuda", lr_net=0.01, epoch_eval_train=1, batch_train=8, distill=False)
    images_train, labels_train = torch.randn(16, din), torch.randn(16, dt)
    loader = _loader(torch.randn(n_img, din), n_img, caps)
    bert = torch.randn(n_img * caps, dt)
    net2, accs, res = epoch.evaluate_synset(0, net, images_train, labels_train, loader, args, bert)
    assert net2 is net and len(accs) == 2 and args.distill is True
    assert list(res.keys()) == list(RR.RESULT_KEYS) and all(0.0 <= v <= 100.0 for v in res.values())
    # the retrieval tail equals the oracle on the trained net's embeddings
    with torch.no_grad():
        net.eval()
        xi = net.image_encoder(torch.cat([b[0] for b in loader]).cuda())
        xi = xi / xi.norm(dim=1, keepdim=True)
        yt = net.text_projection(bert.cuda())
        yt = yt / yt.norm(dim=1, keepdim=True)
        S = (14.285714 * xi @ yt.t()).cpu().numpy()
    ref = RR.itm_eval_ref(S, np.ascontiguousarray(S.T), loader.dataset.txt2img, loader.dataset.img2txt)
    for k in RR.RESULT_KEYS:
        assert abs(res[k] - ref[k]) <= 100.0 / n_img + 1e-9, k


def test_evaluate_synset_with_kernel_backed_clip_model_tracks_torch_training():
    """networks.CLIPModel_full (vldd_clip_loss under autograd) trained by evaluate_synset follows the same trajectory as
    the plain-torch model with identical initial weights: same per-epoch training accuracy, parameters within fp32
    tolerance after the SGD steps, same retrieval result."""
    import copy
    import types
    from multimodal_dataset_distillation_b200 import epoch, networks
    torch.manual_seed(1)
    dt, d, din, n_img, caps = 16, 32, 24, 40, 5
    ref_net = FakeCLIP(dt, d, din)
    ref_net.text_projection.dropout.p = 0.0                       # deterministic: no dropout draws to align
    enc = copy.deepcopy(ref_net.image_encoder)
    net = networks.CLIPModel_full(types.SimpleNamespace(distill=True), image_encoder=enc, image_embedding=d, text_embedding=dt)
    net.text_projection.dropout.p = 0.0
    net.text_projection.load_state_dict(ref_net.text_projection.state_dict())     # same module tree -> same keys
    images_train, labels_train = torch.randn(16, din), torch.randn(16, dt)
    feats, bert = torch.randn(n_img, din), torch.randn(n_img * caps, dt)
    out = []
    for model in (ref_net, net):
        torch.manual_seed(7)                                      # same DataLoader shuffles
        args = types.SimpleNamespace(device="cuda", lr_net=0.05, epoch_eval_train=2, batch_train=8, distill=False)
        out.append(epoch.evaluate_synset(0, model, images_train, labels_train, _loader(feats, n_img, caps), args, bert))
    (_, accs_ref, res_ref), (_, accs_got, res_got) = out
    assert accs_got == accs_ref
    for (n1, p1), (n2, p2) in zip(sorted(ref_net.text_projection.named_parameters()), sorted(net.text_projection.named_parameters())):
        assert n1 == n2 and float((p1 - p2).abs().max()) <= 1e-4 * float(p1.abs().max()) + 1e-6, n1
    assert float((ref_net.image_encoder.weight - net.image_encoder.weight).abs().max()) <= 1e-4
    for k in RR.RESULT_KEYS:
        assert abs(res_got[k] - res_ref[k]) <= 100.0 / n_img + 1e-9, k


def test_epoch_with_amp_scaler_runs_the_kernel_backed_model():
    """Fork signature epoch(e, dataloader, net, optimizer_img, optimizer_txt, args, scaler) (epoch.py:59-98): autocast +
    GradScaler around the kernel-backed CLIPModel_full; the scaled loss reaches our backward as `gout`, parameters move,
    the returned averages are finite."""
    import types
    from multimodal_dataset_distillation_b200 import epoch, networks
    torch.manual_seed(2)
    dt, d, din = 16, 32, 24
    net = networks.CLIPModel_full(types.SimpleNamespace(distill=True), image_encoder=nn.Linear(din, d), image_embedding=d,
                                  text_embedding=dt).cuda()
    before = net.text_projection.fc.weight.detach().clone()
    enc_before = net.image_encoder.weight.detach().clone()
    args = types.SimpleNamespace(device="cuda", distill=True)
    opt_i = torch.optim.SGD(net.image_encoder.parameters(), lr=0.05)
    opt_t = torch.optim.SGD(net.text_projection.parameters(), lr=0.05)
    ds = torch.utils.data.TensorDataset(torch.randn(32, din), torch.randn(32, dt))
    loader = torch.utils.data.DataLoader(ds, batch_size=8)
    loss_avg, acc_avg = epoch.epoch(0, loader, net, opt_i, opt_t, args, scaler=torch.amp.GradScaler("cuda"))
    assert np.isfinite(loss_avg) and 0.0 <= acc_avg <= 8.0
    assert not torch.equal(before, net.text_projection.fc.weight.detach())
    assert not torch.equal(enc_before, net.image_encoder.weight.detach())
