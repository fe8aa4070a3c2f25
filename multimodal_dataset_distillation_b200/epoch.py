"""Drop-in mirror of the reference's epoch.py / epoch_original.py for the retrieval-scoring path.

Kept signatures:
  itm_eval(scores_i2t, scores_t2i, txt2img, img2txt) -> dict with the 9 reference keys     epoch.py:219
  epoch_test(dataloader, model, device, bert_test_embed) -> (np[I,T], np[T,I])             epoch_original.py:68
  evaluate_synset(it_eval, net, images_train, labels_train, testloader, args, bert_test_embed,
                  return_loss=False) -> (net, acc_train_list, val_result)                  epoch.py:348
  epoch(e, dataloader, net, optimizer_img, optimizer_txt, args, scaler=None)               epoch_original.py:20

What runs on the B200 kernels: the text head over the test captions, the similarity GEMM, the top-128/-100 fill
and the ranking / recall.  The image encoder (NFNet) and the train-from-synthetic loop stay ordinary PyTorch
(BASELINE.json: feature extraction is outside the hot path).  No CPU fallback: calls raise without CUDA.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import ops

RESULT_KEYS = ops.RESULT_KEYS


def _result_from_counts(c_img, n_img: int, c_txt, n_txt: int) -> dict:
    tr1, tr5, tr10 = (100.0 * int(c) / n_img for c in c_img)
    ir1, ir5, ir10 = (100.0 * int(c) / n_txt for c in c_txt)
    trm, irm = (tr1 + tr5 + tr10) / 3, (ir1 + ir5 + ir10) / 3
    return {"txt_r1": tr1, "txt_r5": tr5, "txt_r10": tr10, "txt_r_mean": trm,
            "img_r1": ir1, "img_r5": ir5, "img_r10": ir10, "img_r_mean": irm, "r_mean": (trm + irm) / 2}


def ranks_to_result(ranks_i2t: torch.Tensor, ranks_t2i: torch.Tensor) -> dict:
    """Device ranks -> the reference's result dict (one 6-int D2H copy)."""
    c = torch.stack([ops.recall_counts(ranks_i2t), ops.recall_counts(ranks_t2i)]).cpu().numpy()
    return _result_from_counts(c[0], ranks_i2t.numel(), c[1], ranks_t2i.numel())


@torch.no_grad()
def itm_eval(scores_i2t, scores_t2i, txt2img, img2txt, return_ranks: bool = False):
    """Rank / recall@1/5/10 for image->text (``txt_*`` keys) and text->image (``img_*`` keys).

    Accepts what the reference passes (numpy [I,T] / [T,I] and the dataset's dict maps) and also CUDA tensors,
    in which case nothing but six counters leaves the device.  Ties are broken by index (lower index first).
    """
    if isinstance(scores_i2t, torch.Tensor) and scores_i2t.is_cuda:
        n_img, n_txt = scores_i2t.shape
        t2i, ptr, idx = ops.maps_to_arrays(txt2img, img2txt, n_img, n_txt)
        dev = scores_i2t.device
        r1, r2 = ops.ranks_from_scores(scores_i2t.float(), scores_t2i.float(), torch.from_numpy(t2i).to(dev),
                                       torch.from_numpy(ptr).to(dev), torch.from_numpy(idx).to(dev))
        res = ranks_to_result(r1, r2)
        return (res, r1.cpu().numpy(), r2.cpu().numpy()) if return_ranks else res
    s1 = scores_i2t.numpy() if isinstance(scores_i2t, torch.Tensor) else np.asarray(scores_i2t)
    s2 = scores_t2i.numpy() if isinstance(scores_t2i, torch.Tensor) else np.asarray(scores_t2i)
    n_img, n_txt = s1.shape
    t2i, ptr, idx = ops.maps_to_arrays(txt2img, img2txt, n_img, n_txt)
    return ops.itm_eval_host(s1, s2, t2i, ptr, idx, want_ranks=return_ranks)


def text_head_flat_param(text_projection) -> torch.Tensor:
    """Flat parameter vector of a ProjectionHead-like module in ReparamModule order (reparam_module.py:28-51)."""
    if isinstance(getattr(text_projection, "flat_param", None), torch.Tensor):      # ReparamModule-wrapped head
        return text_projection.flat_param.detach()
    mod = text_projection
    return torch.cat([p.detach().reshape(-1) for p in (mod.projection.weight, mod.projection.bias, mod.fc.weight,
                                                       mod.fc.bias, mod.layer_norm.weight, mod.layer_norm.bias)])


@torch.no_grad()
def embed_for_eval(dataloader, model, device, bert_test_embed):
    """epoch_original.py:77-92: normalised text embeddings (CUDA head kernels) and image embeddings (model's encoder)."""
    model.eval()
    dev = torch.device(device)
    theta = text_head_flat_param(model.text_projection).to(dev, torch.float32)
    bert = torch.as_tensor(bert_test_embed).to(dev, torch.float32)
    d = theta.numel()
    dt = bert.shape[1]
    # P = d*dt + d*d + 4d  ->  d
    disc = (dt + 4) ** 2 + 4 * d
    dproj = int(round((-(dt + 4) + disc ** 0.5) / 2))
    text_embeds = ops.proj_head_forward(theta, bert, dproj, normalise=True)
    feats = []
    for image, _img_id in dataloader:
        f = model.image_encoder(image.to(dev)).float()
        feats.append(f / f.norm(dim=1, keepdim=True))
    image_embeds = torch.cat(feats, dim=0)
    image_embeds = image_embeds / image_embeds.norm(dim=1, keepdim=True)      # epoch_original.py:92 (second normalise)
    return image_embeds.contiguous(), text_embeds


@torch.no_grad()
def epoch_test(dataloader, model, device, bert_test_embed):
    """Reference-compatible: returns the two dense top-128/-100-filled score matrices as numpy arrays."""
    start = time.time()
    image_embeds, text_embeds = embed_for_eval(dataloader, model, device, bert_test_embed)
    s_i2t, s_t2i = ops.sim_scores(image_embeds, text_embeds, ops.LOGIT_SCALE_EVAL)
    out = ops.topk_fill(s_i2t, 128, -100.0).cpu().numpy(), ops.topk_fill(s_t2i, 128, -100.0).cpu().numpy()
    print("Evaluation time {:.3f}s".format(time.time() - start))
    return out


@torch.no_grad()
def epoch_test_metrics(dataloader, model, device, bert_test_embed):
    """Embeddings -> result dict without materialising anything on the host (fork epoch.py:103 returns the dict too)."""
    image_embeds, text_embeds = embed_for_eval(dataloader, model, device, bert_test_embed)
    ds = dataloader.dataset
    n_img, n_txt = image_embeds.shape[0], text_embeds.shape[0]
    t2i, ptr, idx = ops.maps_to_arrays(ds.txt2img, ds.img2txt, n_img, n_txt)
    dev = image_embeds.device
    # large galleries: ranking fused into the GEMM epilogue (no [I,T] matrix in HBM); small ones: one GEMM + rank kernels
    rank_fn = ops.sim_rank_fused if (n_img * n_txt >= 50_000_000 and image_embeds.shape[1] % 4 == 0) else ops.sim_rank
    r1, r2 = rank_fn(image_embeds, text_embeds, torch.from_numpy(t2i).to(dev), torch.from_numpy(ptr).to(dev),
                     torch.from_numpy(idx).to(dev), ops.LOGIT_SCALE_EVAL)
    return ranks_to_result(r1, r2)


def epoch(e, dataloader, net, optimizer_img, optimizer_txt, args, scaler=None):
    """One training epoch of a CLIPModel_full-like net (epoch_original.py:20-62; fork adds AMP `scaler`)."""
    net = net.to(args.device)
    net.train()
    loss_avg, acc_avg, num_exp = 0.0, 0.0, 0
    for data in dataloader:
        if getattr(args, "distill", False):
            image, caption = data[:2]
        else:
            image, caption = data[0], data[1]
        image = image.to(args.device)
        n_b = image.shape[0]
        if scaler is not None:
            with torch.autocast("cuda"):
                loss, acc = net(image, caption, e)
        else:
            loss, acc = net(image, caption, e)
        loss_avg += loss.item() * n_b
        acc_avg += acc
        num_exp += n_b
        optimizer_img.zero_grad()
        optimizer_txt.zero_grad()
        if scaler is not None:
            scaler.scale(loss).backward()
            scaler.step(optimizer_img)
            scaler.step(optimizer_txt)
            scaler.update()
        else:
            loss.backward()
            optimizer_img.step()
            optimizer_txt.step()
    return loss_avg / max(num_exp, 1), acc_avg / max(num_exp, 1)


class _NullOptimizer:
    """Stands in for the image optimiser when the image tower has no parameters."""

    def zero_grad(self, *a, **k):
        pass

    def step(self, *a, **k):
        pass


def evaluate_synset(it_eval, net, images_train, labels_train, testloader, args, bert_test_embed, return_loss=False):
    """Train `net` on the synthetic set, then retrieval-evaluate it (epoch_original.py:164-195 / epoch.py:348-397)."""
    net = net.to(args.device)
    images_train = images_train.to(args.device)
    labels_train = labels_train.to(args.device)
    args.distill = True                                           # fork epoch.py:357
    lr = float(args.lr_net)
    Epoch = int(args.epoch_eval_train)
    img_params = list(net.image_encoder.parameters())              # empty when the image tower is frozen embeddings (identity)
    optimizer_img = (torch.optim.SGD(img_params, lr=lr, momentum=0.9, weight_decay=0.0005) if img_params else _NullOptimizer())
    optimizer_txt = torch.optim.SGD(net.text_projection.parameters(), lr=lr, momentum=0.9, weight_decay=0.0005)
    dst_train = torch.utils.data.TensorDataset(images_train, labels_train)
    trainloader = torch.utils.data.DataLoader(dst_train, batch_size=args.batch_train, shuffle=True, num_workers=0)
    acc_train_list, val_result = [], None
    for ep in range(Epoch + 1):
        _loss_train, acc_train = epoch(ep, trainloader, net, optimizer_img, optimizer_txt, args)
        acc_train_list.append(acc_train)
        if ep == Epoch:
            val_result = epoch_test_metrics(testloader, net, args.device, bert_test_embed)
    return net, acc_train_list, val_result
