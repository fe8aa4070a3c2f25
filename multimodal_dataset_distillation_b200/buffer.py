"""Mirror of the reference's buffer.py for the part that feeds the hot path: expert trajectories of the text head.

The reference trains `num_experts` CLIP-style teachers (NFNet image tower + text_projection over frozen BERT embeddings)
with `epoch()` and stores, per expert, the parameter snapshot after every epoch (buffer.py:41-115):

    {buffer_path}/{dataset}/{image_encoder}/{text_encoder}/txt_replay_buffer_{n}.pt
        = list[expert] of list[snapshot] of list[param CPU tensors]           (buffer.py:64-68, 94-95, 104-112)

distill.py later matches student trajectories against these files (distill.py:258-283, 450-476).  This mirror produces the
same files for "Mode A" of the north star: the image tower is frozen (precomputed image-encoder embeddings), so only the
text_projection head is trained -- with the kernel-backed symmetric InfoNCE (vldd_clip_loss through
networks.clip_contrastive_loss: head forward, normalise, logits, loss, top-1 counters and every gradient in one C-ABI call)
and the reference's plain SGD (buffer.py:58-59).  No image trajectory file is written (there are no image parameters).
Every CLI flag of buffer.py:119-160 is accepted with the same name, type and default (tests/golden/cli_buffer.json).
"""
from __future__ import annotations

import argparse
import datetime
import os

import numpy as np
import torch

from . import networks, ops


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Parameter Processing")
    A = p.add_argument
    # ---- buffer.py:120-159, verbatim names / types / defaults ----
    A("--dataset", type=str, default="flickr", choices=["flickr", "coco"], help="dataset")
    A("--num_experts", type=int, default=100, help="training iterations")
    A("--lr_teacher_img", type=float, default=0.1, help="learning rate for updating network parameters")
    A("--lr_teacher_txt", type=float, default=0.1, help="learning rate for updating network parameters")
    A("--batch_train", type=int, default=128, help="batch size for training networks")
    A("--dsa", type=str, default="True", choices=["True", "False"], help="whether to use differentiable Siamese augmentation.")
    A("--dsa_strategy", type=str, default="color_crop_cutout_flip_scale_rotate", help="differentiable Siamese augmentation strategy")
    A("--data_path", type=str, default="./data/Flickr30k/", help="dataset path")
    A("--buffer_path", type=str, default="./buffers", help="buffer path")
    A("--train_epochs", type=int, default=50)
    A("--zca", action="store_true")
    A("--decay", action="store_true")
    A("--mom", type=float, default=0, help="momentum")
    A("--l2", type=float, default=0, help="l2 regularization")
    A("--save_interval", type=int, default=10)
    A("--name", type=str, default=datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S"), help="name of wandb run")
    A("--text_pretrained", type=bool, default=True, help="text_pretrained")
    A("--image_pretrained", type=bool, default=True, help="image_pretrained")
    A("--text_trainable", type=bool, default=False, help="text_trainable")
    A("--image_trainable", type=bool, default=True, help="image_trainable")
    A("--batch_size_train", type=int, default=128, help="batch_size_train")
    A("--batch_size_test", type=int, default=128, help="batch_size_test")
    A("--image_root", type=str, default="./Flickr30k/flickr-image-dataset/flickr30k-images/", help="location of image root")
    A("--ann_root", type=str, default="./Flickr30k/ann_file/", help="location of ann root")
    A("--image_size", type=int, default=224, help="image_size")
    A("--k_test", type=int, default=128, help="k_test")
    A("--load_npy", type=bool, default=False, help="load_npy")
    A("--image_encoder", type=str, default="resnet50", choices=["nfnet", "resnet18_gn", "vit_tiny", "nf_resnet50", "nf_regnet", "resnet50"],
      help="image encoder")
    A("--text_encoder", type=str, default="bert", choices=["bert", "clip"], help="text encoder")
    A("--margin", default=0.2, type=float, help="Rank loss margin.")
    A("--measure", default="cosine", help="Similarity measure used (cosine|order)")
    A("--max_violation", action="store_true", help="Use max instead of sum in the rank loss.")
    A("--only_has_image_projection", type=bool, default=False, help="None")
    A("--grounding", type=bool, default=False, help="None")
    A("--distill", type=bool, default=False, help="if distill")
    # ---- this library ----
    A("--embed_path", type=str, default=None,
      help=".npz with frozen train embeddings image_embed [M,d], text_embed [M,dt] and optionally test_image_embed, test_text_embed")
    A("--synthetic", action="store_true", help="random Flickr-shaped embeddings (no datasets are available offline)")
    A("--num_pairs", type=int, default=2048, help="number of synthetic training pairs")
    A("--seed", type=int, default=0)
    return p


def save_dir_of(args) -> str:
    """buffer.py:27-32."""
    d = os.path.join(args.buffer_path, args.dataset)
    if args.dataset in ("CIFAR10", "CIFAR100") and not args.zca:
        d += "_NO_ZCA"
    return os.path.join(d, args.image_encoder, args.text_encoder)


def train_expert(image_embed: torch.Tensor, text_embed: torch.Tensor, args, generator: torch.Generator, test=None):
    """One teacher (buffer.py:44-101 with a frozen image tower): returns (snapshots, per-epoch (loss, acc[, r_mean]))."""
    dev = image_embed.device
    dt, d = text_embed.shape[1], image_embed.shape[1]
    head = networks.ProjectionHead(dt, d).to(dev).train()
    lr = float(args.lr_teacher_txt)
    opt = torch.optim.SGD(head.parameters(), lr=lr, momentum=args.mom, weight_decay=args.l2)        # buffer.py:59
    snaps = [[p.detach().cpu() for p in head.parameters()]]                                           # buffer.py:67
    lr_schedule = [args.train_epochs // 2 + 1]                                                        # buffer.py:69
    n, bs = image_embed.shape[0], int(args.batch_train)
    log = []
    for e in range(int(args.train_epochs)):
        perm = torch.randperm(n, generator=generator).to(dev)
        loss_sum, acc_sum, seen = 0.0, 0.0, 0
        for i in range(0, n, bs):
            idx = perm[i:i + bs]
            opt.zero_grad()
            mask = head.dropout_mask(idx.numel(), dev)
            loss, top1 = networks.clip_contrastive_loss(head.flat_parameters(), text_embed[idx], image_embed[idx],
                                                        ops.LOGIT_SCALE_EVAL, mask)
            loss.backward()
            opt.step()
            loss_sum += float(loss.detach()) * idx.numel()                                            # epoch.py:84-86
            acc_sum += float(top1.sum()) / 2
            seen += idx.numel()
        entry = [loss_sum / seen, acc_sum / seen]
        if test is not None:                                                                          # buffer.py:73-74
            head.eval()
            with torch.no_grad():
                yn = ops.proj_head_forward(head.flat_parameters().detach(), test["text"], d, normalise=True)
            r1, r2 = ops.sim_rank(test["image"], yn, test["t2i"], test["ptr"], test["idx"], ops.LOGIT_SCALE_EVAL)
            from . import epoch as epoch_mod
            entry.append(epoch_mod.ranks_to_result(r1, r2)["r_mean"])
            head.train()
        log.append(tuple(entry))
        snaps.append([p.detach().cpu() for p in head.parameters()])                                   # buffer.py:94-95
        if e in lr_schedule and args.decay:                                                           # buffer.py:97-102
            lr *= 0.1
            opt = torch.optim.SGD(head.parameters(), lr=lr, momentum=args.mom, weight_decay=args.l2)
    return snaps, log


def main(args):
    if not torch.cuda.is_available():
        raise RuntimeError("buffer needs a CUDA device (sm_100a); there is no CPU path")
    args.device = "cuda"                                                                              # buffer.py:20
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(int(args.seed))
    test = None
    if args.synthetic or args.embed_path is None:
        M, dt, d = int(args.num_pairs), 768, 2304
        img = torch.randn(M, d, generator=g)
        txt = 0.05 * img[:, :dt] + (torch.randn(M, dt, generator=g) * 0.5253 - 0.0094)              # weakly paired
    else:
        z = np.load(args.embed_path)
        img, txt = torch.from_numpy(z["image_embed"]).float(), torch.from_numpy(z["text_embed"]).float()
        if "test_image_embed" in z.files:
            ti, tt = torch.from_numpy(z["test_image_embed"]).float(), torch.from_numpy(z["test_text_embed"]).float()
            caps = tt.shape[0] // ti.shape[0]
            ti = (ti / ti.norm(dim=1, keepdim=True)).to(dev).contiguous()
            test = dict(image=ti, text=tt.to(dev).contiguous(),
                        t2i=(torch.arange(tt.shape[0]) // caps).int().to(dev),
                        ptr=(torch.arange(ti.shape[0] + 1) * caps).int().to(dev), idx=torch.arange(tt.shape[0]).int().to(dev))
    img, txt = img.to(dev).contiguous(), txt.to(dev).contiguous()
    save_dir = save_dir_of(args)
    os.makedirs(save_dir, exist_ok=True)
    written = []
    for it in range(int(args.num_experts)):
        snaps, log = train_expert(img, txt, args, g, test)
        tail = "" if len(log[-1]) < 3 else "\tR@Mean: {:.2f}".format(log[-1][2])
        print("Itr: {}\tEpochs: {}\tTrain Loss: {:.4f}\tTrain Acc: {:.2f}{}".format(it, len(log), log[-1][0], log[-1][1], tail))
        n = 0
        while os.path.exists(os.path.join(save_dir, "txt_replay_buffer_{}.pt".format(n))):           # buffer.py:106-108
            n += 1
        path = os.path.join(save_dir, "txt_replay_buffer_{}.pt".format(n))
        print("Saving {}".format(path))
        torch.save([snaps], path)                                                                     # one expert per file, as the fork does
        written.append(path)
    return written


if __name__ == "__main__":
    main(build_parser().parse_args())
