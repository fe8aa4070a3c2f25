"""Drop-in mirror of the reference's distill.py for the trajectory-matching inner loop (text tower, Mode A).

What is kept from the reference:
  * the CLI: every flag of distill.py:625-679 with the same name, type and default, parse_known_args leniency;
  * the expert-buffer file format  {buffer_path}/txt_replay_buffer_{n}.pt = list[expert] of list[snapshot] of
    list[param tensors] (buffer.py:64-68,94-112; read at distill.py:255-283);
  * the semantics of one outer iteration (distill.py:439-613): pick expert / start_epoch, K unrolled student steps
    on the synthetic pairs, matching loss, backward, three momentum-0.5 SGD updates (image, text, lr).

What is different by design: the whole inner loop + backward is ONE call into the CUDA engine
(``ops.unrolled_match``) wrapped in ``UnrolledMatch`` (a torch.autograd.Function, so ``grand_loss.backward()``
still works); expert snapshots live on the device as one flat [experts, snapshots, P] tensor; the image side is the
image-encoder OUTPUT (frozen-NFNet embeddings, BASELINE.json north_star), not pixels.  Multi-GPU: one process per
GPU, each rank runs a different expert segment and the synthetic-data gradients are all-reduced (NCCL).
"""
from __future__ import annotations

import argparse
import datetime
import glob
import math
import os
import types

import numpy as np
import torch

from . import dist as dist_mod
from . import ops

_BOOL = bool  # the reference uses type=bool (any non-empty string is True); kept for CLI compatibility


def build_parser() -> argparse.ArgumentParser:
    """Same flags / types / defaults as distill.py:625-679."""
    p = argparse.ArgumentParser(description="Parameter Processing")
    A = p.add_argument
    A("--distributed", action="store_true")
    A("--max_files", type=int, default=1)
    A("--dataset", type=str, default="roco", choices=["roco", "coco", "flickr"])   # README.md:52 uses flickr
    A("--num_queries", type=int, default=100)
    A("--lr_img", type=float, default=1000)
    A("--lr_txt", type=float, default=1000)
    A("--lr_lr", type=float, default=1e-03)
    A("--Iteration", type=int, default=50000)
    A("--eval_it", type=int, default=50)
    A("--num_eval", type=int, default=5)
    A("--epoch_eval_train", type=int, default=1)
    A("--syn_steps", type=int, default=20)
    A("--mini_batch_size", type=int, default=100)
    A("--max_start_epoch", type=int, default=25)
    A("--expert_epochs", type=int, default=3)
    A("--ipc", type=int, default=1)
    A("--force_save", action="store_true")
    A("--draw", type=_BOOL, default=True)
    A("--transfer", type=_BOOL, default=False)
    A("--std", type=_BOOL, default=False)
    A("--disable_wandb", action="store_true")
    A("--num_experts", type=int, default=100)
    A("--lr_teacher_img", type=float, default=0.1)
    A("--lr_teacher_txt", type=float, default=0.1)
    A("--batch_train", type=int, default=128)
    A("--dsa", type=str, default="True", choices=["True", "False"])
    A("--dsa_strategy", type=str, default="color_crop_cutout_flip_scale_rotate")
    A("--data_path", type=str, default="/kaggle/input/roco-dataset/")
    A("--buffer_path", type=str, default="/kaggle/working")
    A("--train_epochs", type=int, default=50)
    A("--zca", action="store_true")
    A("--decay", action="store_true")
    A("--mom", type=float, default=0)
    A("--l2", type=float, default=0)
    A("--save_interval", type=int, default=10)
    A("--name", type=str, default=datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S"))
    A("--text_pretrained", type=_BOOL, default=True)
    A("--image_pretrained", type=_BOOL, default=True)
    A("--text_trainable", type=_BOOL, default=False)
    A("--image_trainable", type=_BOOL, default=True)
    A("--batch_size_train", type=int, default=128)
    A("--batch_size_test", type=int, default=128)
    A("--image_root", type=str, default="/kaggle/input/roco-dataset/all_data/train/radiology/images/")
    A("--ann_root", type=str, default="/kaggle/input/roco-dataset/all_data/train/radiologytraindata.csv")
    A("--image_size", type=int, default=224)
    A("--k_test", type=int, default=128)
    A("--load_npy", type=_BOOL, default=False)
    A("--image_encoder", type=str, default="resnet50",
      choices=["nfnet", "resnet18_gn", "vit_tiny", "nf_resnet50", "nf_regnet", "resnet50"])
    A("--text_encoder", type=str, default="bert", choices=["bert", "clip"])
    A("--margin", default=0.2, type=float)
    A("--measure", default="cosine")
    A("--max_violation", action="store_true")
    A("--only_has_image_projection", type=_BOOL, default=False)
    A("--grounding", type=_BOOL, default=False)
    A("--distill", type=_BOOL, default=False)
    # additions of this implementation (not in the reference)
    A("--logit_scale_mode", type=str, default="fork", choices=["fork", "upstream"],
      help="fork: logits scaled by the learnable syn_lr_img (distill.py:548); upstream: fixed log(1/0.07) "
           "(distill_original.py:103,430)")
    A("--embed_path", type=str, default=None, help=".npz with image_embed [M,d] and text_embed [M,dt] (frozen encoders)")
    A("--synthetic", action="store_true", help="run on synthetic Flickr-shaped embeddings and experts")
    A("--seed", type=int, default=0)
    A("--segments_in_flight", type=int, default=1,
      help="expert segments per outer step processed concurrently on this GPU (throughput mode, DistillEngine.segments_step); "
           "1 = the reference's one segment per iteration")
    A("--grad_reduce", type=str, default="sum", choices=["sum", "mean"],
      help="how the synthetic-data gradients of the segments of one outer step (segments_in_flight x world size) are combined "
           "before the SGD step: sum (a G-segment step is G times the reference's single-segment step; scale lr_img / lr_txt / "
           "lr_lr down accordingly) or mean (same step size as the reference, lower variance)")
    A("--save_path", type=str, default=None,
      help="directory for the distilled set (distilled_{it}.pt: U, Y, syn_lr_img, syn_lr_txt[, sentences]) at the evaluation "
           "iterations and at exit; default {buffer_path}/distilled")
    A("--student_dropout", type=float, default=0.1,
      help="dropout of the student text_projection during the unroll (networks.py:629,636; students are in train mode, "
           "distill.py:446-447); 0 gives the deterministic parity mode")
    return p


def parse_args(argv=None):
    args, unknown = build_parser().parse_known_args(argv)      # distill.py:680-682
    if unknown:
        print("Warning: Ignoring unknown arguments:", unknown)
    return args


# ------------------------------------------------------------------------------------------------------
class UnrolledMatch(torch.autograd.Function):
    """loss = |theta_K - theta_tgt|^2 / |theta_0 - theta_tgt|^2 after K unrolled steps (distill.py:509-598).

    Differentiable w.r.t. text_syn (Y), image embeddings (U), syn_lr and the logit scale; `.backward()` replays
    nothing -- the engine already produced the gradients in the same call (reverse sweep).
    """

    @staticmethod
    def forward(ctx, Y, U, lr, scale, theta0, theta_tgt, perms, masks, workspace):
        res = ops.unrolled_match(theta0, theta_tgt, Y.detach(), U.detach(), lr, scale, perms, masks, workspace)
        ctx.save_for_backward(res["dY"], res["dU"], res["out5"])
        ctx.lr_shape, ctx.scale_shape = lr.shape, scale.shape
        ctx.aux = res
        return res["out5"][2].clone()

    @staticmethod
    def backward(ctx, gout):
        dY, dU, out5 = ctx.saved_tensors
        return (gout * dY, gout * dU, (gout * out5[3]).reshape(ctx.lr_shape), (gout * out5[4]).reshape(ctx.scale_shape),
                None, None, None, None, None)


def nearest_neighbor(sentences, query_embeddings, database_embeddings):
    """distill.py:89-95: for every synthetic text embedding the training caption whose embedding is most cosine-similar.

    The reference loops over the queries on the CPU (sklearn cosine_similarity against all ~145k captions per query);
    here both sets are normalised and compared with one tensor-core GEMM + a row arg-max on the device.  Accepts numpy
    arrays or tensors on any device, returns a list of sentences like the reference.
    """
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    if dev is None:
        raise RuntimeError("nearest_neighbor needs a CUDA device (this library has no CPU path)")
    q = torch.as_tensor(query_embeddings).to(dev, torch.float32)
    b = torch.as_tensor(database_embeddings).to(dev, torch.float32)
    if q.dim() == 1:
        q = q.reshape(1, -1)
    idx = ops.nearest_rows(q.contiguous(), b.contiguous()).cpu().tolist()
    return [sentences[i] for i in idx]


def flatten_snapshot(params, device=None) -> torch.Tensor:
    """torch.cat([p.reshape(-1) ...]) of one snapshot (distill.py:471-476), done once per buffer instead of per iteration."""
    flat = torch.cat([torch.as_tensor(p).reshape(-1).float() for p in params], 0)
    return flat.to(device) if device is not None else flat


def load_expert_buffers(buffer_path: str, kind: str = "txt", max_files: int | None = None, device="cuda") -> torch.Tensor:
    """Reads {kind}_replay_buffer_{n}.pt files (buffer.py:104-112) into one [experts, snapshots, P] device tensor."""
    files = sorted(glob.glob(os.path.join(buffer_path, f"{kind}_replay_buffer_*.pt")),
                   key=lambda f: int(os.path.splitext(f)[0].rsplit("_", 1)[1]))
    if not files:
        raise AssertionError("No buffers detected at {}".format(buffer_path))       # distill_original.py:183-184
    if max_files:
        files = files[:max_files]
    experts = []
    for f in files:
        for traj in torch.load(f, map_location="cpu"):
            experts.append(torch.stack([flatten_snapshot(snap) for snap in traj]))
    return torch.stack(experts).to(device)


class DistillEngine:
    """State of the distillation: synthetic pairs, learnable student lr(s), outer momentum-SGD, expert segments."""

    def __init__(self, image_embed: torch.Tensor, text_embed: torch.Tensor, experts: torch.Tensor, args,
                 device="cuda", process_group=None, rank: int = 0, world: int = 1):
        self.args = args
        self.dev = torch.device(device)
        self.U = image_embed.to(self.dev, torch.float32).contiguous().requires_grad_(True)       # "image_syn" (Mode A)
        self.Y = text_embed.to(self.dev, torch.float32).contiguous().requires_grad_(True)        # text_syn  distill.py:231
        self.syn_lr_img = torch.tensor(float(args.lr_teacher_img), device=self.dev, requires_grad=True)   # distill.py:235
        self.syn_lr_txt = torch.tensor(float(args.lr_teacher_txt), device=self.dev, requires_grad=True)   # distill.py:236
        self.experts = experts                      # [E, S, P] on device
        self.N, self.dt = self.Y.shape
        self.d = self.U.shape[1]
        self.K = int(args.syn_steps)
        self.B = min(int(args.mini_batch_size), self.N)
        self.ws = ops.UnrollWorkspace(self.N, self.B, self.K, self.dt, self.d, self.dev)
        # momentum buffers of the outer SGD, one allocation laid out like UnrollWorkspace.pack: [U | Y | lr_img, lr_txt | pad]
        n_u, n_y = self.N * self.d, self.N * self.dt
        self.buf_pack = torch.zeros(n_u + n_y + 8, device=self.dev)
        self.bufs = {"U": self.buf_pack[:n_u].view(self.N, self.d), "Y": self.buf_pack[n_u:n_u + n_y].view(self.N, self.dt)}
        self.buf_lr = self.buf_pack[n_u + n_y:n_u + n_y + 2]
        self.first = True
        self._perm_io = None                                 # lazily created upload path of host-drawn minibatch indices (step_fast)
        self.pg = process_group
        self._init_sampling(int(getattr(args, "seed", 0)), rank, world, int(experts.shape[0]))
        self.rng_state = ops.make_rng_state(self.seed_base + 7919 * (self.rank + 1), self.dev)
        self.fixed_scale = torch.tensor(ops.LOGIT_SCALE_UPSTREAM, device=self.dev)
        self._lanes = []                            # (stream, workspace) pairs of segments_step

    def _init_sampling(self, seed: int, rank: int, world: int, n_experts: int):
        """Host-side sampling state (no device work).  Every rank must sample DIFFERENT segments, minibatches and dropout
        masks -- with one seed everywhere the all-reduce would only multiply one gradient by the world size -- so the host
        generator, the device Philox seed and the expert cursor are offset by the rank; the cursor then advances by the
        world size, i.e. ranks walk disjoint residue classes of the expert list (distill.py:450-465 consumes them in order)."""
        self.rank, self.world = int(rank), max(int(world), 1)
        self.seed_base = int(seed) * 1000003
        self.gen = torch.Generator().manual_seed(self.seed_base + self.rank)
        self.expert_idx = self.rank % max(int(n_experts), 1)

    # -- one expert segment -> loss and grads (distill.py:466-606) --
    def segment_loss(self, expert: int, start_epoch: int, perms: torch.Tensor | None = None, masks=None, workspace=None):
        a = self.args
        ws = self.ws if workspace is None else workspace
        theta0 = self.experts[expert, start_epoch]
        theta_tgt = self.experts[expert, start_epoch + int(a.expert_epochs)]
        if perms is None:
            perms = self.draw_perms()
        perms = perms.to(self.dev)
        fork = getattr(a, "logit_scale_mode", "fork") == "fork"
        scale = self.syn_lr_img if fork else self.fixed_scale
        p_drop = float(getattr(a, "student_dropout", 0.0))
        if masks is None and p_drop > 0.0 and self.K > 0:
            masks = ops.fill_dropout_masks(ws, p_drop, self.rng_state)         # fresh masks per call, as nn.Dropout does
        return UnrolledMatch.apply(self.Y, self.U, self.syn_lr_txt, scale, theta0, theta_tgt, perms, masks, ws)

    def draw_perms(self) -> torch.Tensor:
        """distill.py:510-511: a fresh randperm(N)[:B] per student step, from this rank's host generator."""
        return torch.stack([torch.randperm(self.N, generator=self.gen)[: self.B] for _ in range(self.K)])

    def _grad_scale(self, segments_here: int = 1) -> float:
        if getattr(self.args, "grad_reduce", "sum") != "mean":
            return 1.0
        world = self.world
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size(self.pg)
        return 1.0 / float(segments_here * world)

    def step_fast(self, expert: int | None = None, start_epoch: int | None = None, perms: torch.Tensor | None = None,
                  theta0: torch.Tensor | None = None, theta_tgt: torch.Tensor | None = None):
        """One whole outer iteration (distill.py:439-613) without autograd in the loop: ONE engine call (which also draws
        the dropout masks), one all-reduce of the packed gradient buffer when a process group is up, ONE update kernel.
        No torch kernel is launched.  Returns the loss as a view of the workspace (valid until the next call); a
        non-finite loss leaves the synthetic set untouched and raises ``ws.skipped`` (distill.py:599-600).
        theta0 / theta_tgt override the resident experts (host-streamed segments, SegmentPrefetcher)."""
        a, ws = self.args, self.ws
        if theta0 is None:
            theta0 = self.experts[expert, start_epoch]
            theta_tgt = self.experts[expert, start_epoch + int(a.expert_epochs)]
        perm_slot = None
        if perms is None:
            # drawn on the host in the reference's order (distill.py:510-511) and uploaded on a side stream into one of two device
            # slots: the compute stream only waits for an event -- a copy-engine operation IN the compute stream costs 10-20 us of
            # engine switches per iteration (and queues behind segment uploads, see StepIO)
            if self._perm_io is None:
                self._perm_io = dict(stream=torch.cuda.Stream(device=self.dev), n=0,
                                     slots=[torch.empty(max(self.K, 1), self.B, dtype=torch.int64, device=self.dev) for _ in range(2)],
                                     up=[torch.cuda.Event() for _ in range(2)], free=[torch.cuda.Event() for _ in range(2)])
                for e in self._perm_io["free"]:
                    e.record(torch.cuda.current_stream(self.dev))
            io = self._perm_io
            perm_slot = io["n"] % 2
            io["n"] += 1
            host = self.draw_perms().pin_memory()
            with torch.cuda.stream(io["stream"]):
                io["stream"].wait_event(io["free"][perm_slot])              # the call two iterations ago has gathered from it
                io["slots"][perm_slot].copy_(host, non_blocking=True)
                io["up"][perm_slot].record(io["stream"])
            torch.cuda.current_stream(self.dev).wait_event(io["up"][perm_slot])
            perms = io["slots"][perm_slot]
        fork = getattr(a, "logit_scale_mode", "fork") == "fork"
        scale = self.syn_lr_img if fork else self.fixed_scale
        p_drop = float(getattr(a, "student_dropout", 0.0)) if self.K > 0 else 0.0
        ops.unrolled_match(theta0, theta_tgt, self.Y.detach(), self.U.detach(), self.syn_lr_txt.detach(), scale.detach(), perms,
                           None, ws, dropout_p=p_drop, rng_state=self.rng_state if p_drop > 0 else None, clone_results=False)
        if perm_slot is not None:
            self._perm_io["free"][perm_slot].record(torch.cuda.current_stream(self.dev))
        if self.pg is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            if torch.distributed.get_world_size(self.pg) > 1:
                torch.distributed.all_reduce(ws.pack, group=self.pg)     # [dU | dY | out5]: 1.23 MB at Flickr shape
        self.apply_update(ws)
        return ws.out5[2]

    def apply_update(self, ws, segments_here: int = 1):
        """Fused momentum-SGD update of (U, Y, syn_lr_img, syn_lr_txt) from a workspace's packed gradients."""
        a = self.args
        n_u, n_y = self.N * self.d, self.N * self.dt
        fork = getattr(a, "logit_scale_mode", "fork") == "fork"
        ops.outer_update(self.U.detach(), ws.dU, self.buf_pack[:n_u], float(a.lr_img),
                         self.Y.detach(), ws.dY, self.buf_pack[n_u:n_u + n_y], float(a.lr_txt),
                         self.syn_lr_img.detach(), self.syn_lr_txt.detach(), ws.out5[4:5] if fork else None, ws.out5[3:4],
                         self.buf_pack[n_u + n_y:n_u + n_y + 2], float(a.lr_lr), 0.5, self.first,
                         self._grad_scale(segments_here), ws.out5[2:3], ws.skipped)
        self.first = False

    def segments_step(self, segments, perms_list=None):
        """Throughput mode: several expert segments of ONE outer step in flight at once on this GPU.

        One segment is a chain of ~290 dependent kernels of 3-20 us, so a single segment leaves the machine waiting on
        pipeline fills and drains; independent segments (one CUDA stream and one engine workspace each) fill those gaps:
        measured 505 -> 626 (2 in flight) -> 662 (3) segment-iterations/s on one B200, 4 is slower again
        (profiles/concurrent_segments.py).  The gradients are summed before the outer update, i.e. the same G-segment
        minibatch the multi-GPU run computes (and composes with it: every rank may keep several segments in flight).
        `segments`: list of (expert, start_epoch).  Returns the list of losses.
        """
        main = torch.cuda.current_stream(self.dev)
        while len(self._lanes) < len(segments):
            self._lanes.append((torch.cuda.Stream(device=self.dev),
                                ops.UnrollWorkspace(self.N, self.B, self.K, self.dt, self.d, self.dev)))
        losses = []
        for j, (e, s) in enumerate(segments):
            st, ws = self._lanes[j]
            st.wait_stream(main)
            with torch.cuda.stream(st):
                losses.append(self.segment_loss(e, s, None if perms_list is None else perms_list[j], workspace=ws))
        for j in range(len(segments)):
            main.wait_stream(self._lanes[j][0])
        total = losses[0]
        for l in losses[1:]:
            total = total + l
        self.outer_step(total)
        return losses

    def sample_segment(self):
        """distill.py:450-470: experts are consumed in order, start_epoch ~ U{0..max_start_epoch-1}."""
        e = self.expert_idx
        self.expert_idx = (self.expert_idx + self.world) % self.experts.shape[0]     # ranks walk disjoint residue classes
        hi = min(int(self.args.max_start_epoch), self.experts.shape[1] - int(self.args.expert_epochs))
        s = int(torch.randint(0, max(hi, 1), (1,), generator=self.gen))
        return e, s

    def outer_step(self, loss: torch.Tensor):
        """zero_grad / backward / all-reduce / three SGD(momentum=0.5) steps (distill.py:603-613)."""
        for t in (self.U, self.Y, self.syn_lr_img, self.syn_lr_txt):
            t.grad = None
        loss.backward()
        g_lr = torch.stack([self.syn_lr_img.grad if self.syn_lr_img.grad is not None else torch.zeros((), device=self.dev),
                            self.syn_lr_txt.grad])
        if self.pg is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            # one NCCL all-reduce of the packed [dU | dY | dlr_img, dlr_txt] buffer (1.23 MB at Flickr shape)
            dist_mod.allreduce_packed([self.U.grad, self.Y.grad, g_lr], group=self.pg)
        a = self.args
        gs = self._grad_scale(max(int(getattr(a, "segments_in_flight", 1)), 1))
        if gs != 1.0:                                                # --grad_reduce mean
            self.U.grad.mul_(gs); self.Y.grad.mul_(gs); g_lr = g_lr * gs
        with torch.no_grad():
            ops.momentum_sgd_(self.U, self.U.grad, self.bufs["U"], float(a.lr_img), 0.5, self.first)
            ops.momentum_sgd_(self.Y, self.Y.grad, self.bufs["Y"], float(a.lr_txt), 0.5, self.first)
            lrs = torch.stack([self.syn_lr_img.detach(), self.syn_lr_txt.detach()])
            ops.momentum_sgd_(lrs, g_lr.contiguous(), self.buf_lr, float(a.lr_lr), 0.5, self.first)
            self.syn_lr_img.copy_(lrs[0])
            self.syn_lr_txt.copy_(lrs[1])
        self.first = False

    def iteration(self):
        g = int(getattr(self.args, "segments_in_flight", 1))
        if g > 1:
            return self.segments_step([self.sample_segment() for _ in range(g)])[0]
        e, s = self.sample_segment()
        return self.step_fast(e, s)


class SegmentPrefetcher:
    """Host-resident expert trajectories (the reference keeps them as CPU tensors and uploads theta_start / theta_target
    every iteration, distill.py:466-476) streamed to the device one iteration ahead.

    Two device slots; the copy of the NEXT segment runs on its own stream from pinned memory while the engine works
    on the current one, so the 2 x 28 MB upload is hidden behind compute instead of sitting in front of it.
    """

    def __init__(self, experts_host: torch.Tensor, device="cuda"):
        self.host = experts_host if experts_host.is_pinned() else experts_host.pin_memory()
        self.dev = torch.device(device)
        P = self.host.shape[-1]
        self.slots = [dict(th0=torch.empty(P, device=self.dev), tgt=torch.empty(P, device=self.dev),
                           ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.n = 0
        self.pending = None
        for sl in self.slots:
            sl["free"].record(torch.cuda.current_stream(self.dev))

    def prefetch(self, expert: int, start_epoch: int, expert_epochs: int):
        sl = self.slots[self.n % 2]
        self.n += 1
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(sl["free"])                  # the engine has staged the slot's previous content
            sl["th0"].copy_(self.host[expert, start_epoch], non_blocking=True)
            sl["tgt"].copy_(self.host[expert, start_epoch + expert_epochs], non_blocking=True)
            sl["ready"].record(self.copy_stream)
        self.pending = sl

    def get(self):
        """Buffers of the prefetched segment, valid on the current stream; call release() after the engine call."""
        sl = self.pending
        torch.cuda.current_stream(self.dev).wait_event(sl["ready"])
        return sl

    def release(self, sl):
        sl["free"].record(torch.cuda.current_stream(self.dev))


class LossRing:
    """Every step's loss on the host without a copy-engine operation in the compute stream: a KERNEL copies the 4 bytes into a
    small device ring, a stream of its own fetches them; the host reads step i while later steps are already enqueued."""

    def __init__(self, device, ring: int = 4):
        self.dev = torch.device(device)
        self.n_ring = ring
        self.ring = torch.zeros(ring, dtype=torch.float32, device=self.dev)
        self.host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(ring)]
        self.staged = [torch.cuda.Event() for _ in range(ring)]
        self.down = [torch.cuda.Event() for _ in range(ring)]
        self.d2h_stream = torch.cuda.Stream(device=self.dev)

    def push(self, i: int, loss: torch.Tensor):
        r = i % self.n_ring
        torch.mul(loss.detach().reshape(1), 1.0, out=self.ring[r:r + 1])       # a kernel, not a copy-engine operation
        self.staged[r].record(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.d2h_stream):
            self.d2h_stream.wait_event(self.staged[r])
            self.host[r].copy_(self.ring[r], non_blocking=True)
            self.down[r].record(self.d2h_stream)

    def read(self, i: int) -> float:
        """Loss of step i (blocks until its 4 bytes have arrived; i must be within the last `ring` pushed steps)."""
        r = i % self.n_ring
        self.down[r].synchronize()
        return float(self.host[r])


class StepIO:
    """The small per-step host <-> device traffic of the fast loop, kept OUT of the compute stream.

    With host-resident trajectories a 57 MB segment upload is in flight on the copy stream most of the time.  A copy-engine
    operation enqueued on the compute stream -- the 6 KB of minibatch indices going up, the 4-byte loss coming down -- can be
    queued behind that upload and then holds up every kernel behind it (measured: 2.56 instead of 1.53 ms per step as soon as
    the host runs one step ahead).  So: indices go up on the prefetcher's copy stream into one of `depth` device slots (the
    engine accepts any address, see ops.unrolled_match), and the loss is copied by a KERNEL into a small device ring and
    fetched from there on a stream of its own.  The host can then read every step's loss one step late and never stalls the GPU.
    """

    def __init__(self, K: int, B: int, device, copy_stream: torch.cuda.Stream, depth: int = 2, ring: int = 4):
        self.dev = torch.device(device)
        self.depth, self.n_ring = depth, ring
        self.copy_stream = copy_stream
        self.perms = [torch.empty(max(K, 1), B, dtype=torch.int64, device=self.dev) for _ in range(depth)]
        self.up = [torch.cuda.Event() for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        for e in self.done:
            e.record(torch.cuda.current_stream(self.dev))
        self.losses = LossRing(self.dev, ring)

    def upload_perms(self, i: int, perms_host: torch.Tensor):
        """Minibatch indices of step i (pinned host tensor [K, B] int64) -> device slot, on the copy stream."""
        k = i % self.depth
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.done[k])            # step i - depth has gathered from this slot
            self.perms[k].copy_(perms_host, non_blocking=True)
            self.up[k].record(self.copy_stream)

    def perms_for(self, i: int) -> torch.Tensor:
        k = i % self.depth
        torch.cuda.current_stream(self.dev).wait_event(self.up[k])
        return self.perms[k]

    def step_done(self, i: int, loss: torch.Tensor):
        """Call right after step i was enqueued: frees its index slot and sends its loss towards the host."""
        self.done[i % self.depth].record(torch.cuda.current_stream(self.dev))
        self.losses.push(i, loss)

    def loss(self, i: int) -> float:
        """Loss of step i on the host (blocks until its 4 bytes have arrived; i must be within the last `ring` steps)."""
        return self.losses.read(i)


class SegmentCache:
    """Host-resident expert trajectories with a device-side LRU of uploaded snapshots.

    The reference uploads theta_start and theta_target from CPU lists every iteration (distill.py:466-476) although the
    same (expert, epoch) snapshots come back again and again (experts are cycled, start epochs are drawn from a small range,
    and one iteration's target snapshot is a later iteration's start).  When all trajectories fit in HBM keep them resident
    (``load_expert_buffers``); when they do not, this cache keeps the most recently used `capacity` snapshots on the device,
    uploads only misses -- on a copy stream, one iteration ahead, like SegmentPrefetcher -- and evicts least-recently-used
    slots.  Same prefetch / get / release protocol as SegmentPrefetcher; ``h2d_bytes`` counts what was actually copied.
    """

    def __init__(self, experts_host: torch.Tensor, device="cuda", capacity: int = 64):
        import collections
        self.host = experts_host if experts_host.is_pinned() else experts_host.pin_memory()
        self.dev = torch.device(device)
        self.P = int(self.host.shape[-1])
        self.capacity = max(int(capacity), 4)            # two snapshots in use + two being prefetched
        self.slots = collections.OrderedDict()           # (expert, snapshot) -> dict(buf, ready, free); order = recency
        self.spare = []
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.pending = None
        self.h2d_bytes = 0
        self.hits = self.misses = 0

    def _slot_for(self, key):
        sl = self.slots.get(key)
        if sl is not None:
            self.slots.move_to_end(key)
            self.hits += 1
            return sl
        self.misses += 1
        if len(self.slots) < self.capacity:
            sl = dict(buf=torch.empty(self.P, device=self.dev), ready=torch.cuda.Event(), free=torch.cuda.Event())
            sl["free"].record(torch.cuda.current_stream(self.dev))
        else:
            _, sl = self.slots.popitem(last=False)       # least recently used
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(sl["free"])      # the engine has staged this slot's previous content
            sl["buf"].copy_(self.host[key[0], key[1]], non_blocking=True)
            sl["ready"].record(self.copy_stream)
        self.h2d_bytes += self.P * 4
        self.slots[key] = sl
        return sl

    def prefetch(self, expert: int, start_epoch: int, expert_epochs: int):
        a = self._slot_for((int(expert), int(start_epoch)))
        b = self._slot_for((int(expert), int(start_epoch) + int(expert_epochs)))
        self.pending = dict(th0=a["buf"], tgt=b["buf"], _slots=(a, b))

    def get(self):
        sl = self.pending
        st = torch.cuda.current_stream(self.dev)
        for x in sl["_slots"]:
            st.wait_event(x["ready"])
        return sl

    def release(self, sl):
        st = torch.cuda.current_stream(self.dev)
        for x in sl["_slots"]:
            x["free"].record(st)


def synthetic_experts(n_experts: int, n_snapshots: int, dt: int, d: int, seed: int = 0, step: float = 0.01) -> torch.Tensor:
    """Random-walk expert trajectories of ProjectionHead shape (no checkpoints are available offline)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_experts):
        b1, b2 = 1.0 / math.sqrt(dt), 1.0 / math.sqrt(d)
        th = torch.cat([(torch.rand(d * dt, generator=g) * 2 - 1) * b1, (torch.rand(d, generator=g) * 2 - 1) * b1,
                        (torch.rand(d * d, generator=g) * 2 - 1) * b2, (torch.rand(d, generator=g) * 2 - 1) * b2,
                        torch.ones(d), torch.zeros(d)])
        snaps = [th]
        for _ in range(n_snapshots - 1):
            snaps.append(snaps[-1] + step * torch.randn(th.shape, generator=g))
        out.append(torch.stack(snaps))
    return torch.stack(out)


class _EmbeddingLoader(list):
    """Test "dataloader" over precomputed image-encoder embeddings: yields (features, ids) batches and carries the
    dataset maps that epoch_test / itm_eval read (flickr30k_dataset.py:110-118)."""


def _load_test_split(args, dev):
    """Optional test embeddings next to the training ones: test_image_embed [I,d], test_text_embed [T,dt] (T = C * I)."""
    if getattr(args, "embed_path", None) is None or getattr(args, "synthetic", False):
        return None
    z = np.load(args.embed_path)
    if "test_image_embed" not in z.files:
        return None
    ti, tt = torch.from_numpy(z["test_image_embed"]).float(), torch.from_numpy(z["test_text_embed"]).float()
    caps = tt.shape[0] // ti.shape[0]
    loader = _EmbeddingLoader((ti[i:i + 256], torch.arange(i, min(i + 256, ti.shape[0]))) for i in range(0, ti.shape[0], 256))
    loader.dataset = types.SimpleNamespace(txt2img={t: t // caps for t in range(tt.shape[0])},
                                           img2txt={i: list(range(caps * i, caps * i + caps)) for i in range(ti.shape[0])})
    return dict(loader=loader, bert=tt)


def evaluate_synthetic_set(eng: "DistillEngine", test, args):
    """distill.py:293-330: `num_eval` fresh models trained on the current synthetic set (evaluate_synset, lr_net = the
    learned syn_lr_img), each retrieval-evaluated on the test split; returns the list of result dicts.  Mode A: the
    synthetic "images" are image-encoder embeddings, so the evaluated model's image tower is the identity."""
    from . import epoch as epoch_mod, networks
    results = []
    for it_eval in range(int(args.num_eval)):
        net_eval = networks.CLIPModel_full(args, image_encoder=torch.nn.Identity(), image_embedding=eng.d, text_embedding=eng.dt)
        ev = types.SimpleNamespace(device=str(eng.dev), lr_net=float(eng.syn_lr_img.detach()), epoch_eval_train=int(args.epoch_eval_train),
                                   batch_train=int(args.batch_train), distill=True)
        _, _acc, val = epoch_mod.evaluate_synset(it_eval, net_eval, eng.U.detach().clone(), eng.Y.detach().clone(), test["loader"],
                                                 ev, test["bert"])
        print("Evaluate_%02d: Img R@1 = %.4f, Img R@5 = %.4f, Img R@10 = %.4f, Img R@Mean = %.4f, Txt R@1 = %.4f, Txt R@5 = %.4f, "
              "Txt R@10 = %.4f, Txt R@Mean = %.4f, R@Mean = %.4f" % (it_eval, val["img_r1"], val["img_r5"], val["img_r10"],
                                                                     val["img_r_mean"], val["txt_r1"], val["txt_r5"], val["txt_r10"],
                                                                     val["txt_r_mean"], val["r_mean"]))
        results.append(val)
    return results


def save_distilled(eng: "DistillEngine", path: str, it: int, sentences=None) -> str:
    """The distilled set as the reference keeps it around for evaluation / visualisation (distill.py:331-384): synthetic
    image embeddings, synthetic text embeddings, the learned student learning rates, and the nearest training captions."""
    os.makedirs(path, exist_ok=True)
    out = os.path.join(path, f"distilled_{it}.pt")
    blob = dict(iteration=it, U=eng.U.detach().cpu(), Y=eng.Y.detach().cpu(), syn_lr_img=float(eng.syn_lr_img.detach()),
                syn_lr_txt=float(eng.syn_lr_txt.detach()))
    if sentences is not None:
        blob["sentences"] = sentences
    torch.save(blob, out)
    return out


def main(args):
    if not torch.cuda.is_available():
        raise RuntimeError("distill needs a CUDA device (sm_100a); there is no CPU path")
    # one process per GPU (torchrun): every rank runs a different expert segment per outer step against the same
    # replicated synthetic set; gradients are all-reduced (NCCL) before the identical update on every rank
    rank, world, local = dist_mod.init_from_env()
    torch.cuda.set_device(local)
    if world > 1:
        dist_mod.bind_to_local_numa(local)                                   # host-side staging next to this rank's GPU
    dev = torch.device("cuda", local)
    args.device = str(dev)                                                   # distill.py:214
    N = int(args.num_queries)
    train_text, sentences = None, None
    if args.synthetic or args.embed_path is None:
        g = torch.Generator().manual_seed(args.seed)                         # same seed on every rank: replicated synthetic set
        dt, d = 768, 2304
        img = torch.randn(N, d, generator=g)
        txt = torch.randn(N, dt, generator=g) * 0.5253 - 0.0094             # distill_original.py:147
        experts = synthetic_experts(max(2, world), int(args.max_start_epoch) + int(args.expert_epochs) + 1, dt, d, args.seed).to(dev)
    else:
        z = np.load(args.embed_path, allow_pickle=True)
        sel = np.random.default_rng(args.seed).permutation(len(z["image_embed"]))[:N]   # distill.py:231 random real pairs
        img, txt = torch.from_numpy(z["image_embed"][sel]), torch.from_numpy(z["text_embed"][sel])
        experts = load_expert_buffers(args.buffer_path, "txt", args.max_files, dev)
        if "sentences" in z.files:                                           # for the nearest-caption read-out (distill.py:374)
            train_text, sentences = torch.from_numpy(z["text_embed"]).float(), [str(x) for x in z["sentences"]]
    eng = DistillEngine(img, txt, experts, args, dev, rank=rank, world=world)
    test = _load_test_split(args, dev)
    eval_it_pool = list(range(0, int(args.Iteration) + 1, max(int(args.eval_it), 1))) if test is not None else []   # distill.py:285
    save_path = args.save_path or os.path.join(str(args.buffer_path), "distilled")
    eng.eval_history = []
    if rank == 0:
        os.makedirs(save_path, exist_ok=True)                                # fail now, not after the last iteration
    # distill.py:599-600: a NaN loss ends the run BEFORE the update is applied.  The update kernel itself refuses to step on a
    # non-finite loss (device-side, every iteration); the host reads every iteration's loss too, one iteration late, so that
    # enqueueing iteration i+1 overlaps the GPU's work on iteration i instead of waiting for a 4-byte result.
    losses = LossRing(dev)
    stamp = lambda: datetime.datetime.now().strftime("[%Y-%m-%d %H:%M:%S]")

    def read_loss(j):                       # loss of iteration j; True when the run has to stop
        v = losses.read(j)
        if math.isnan(v) or math.isinf(v):
            if rank == 0:
                print("%s iter = %04d, loss = %s: stopping, synthetic set left at the last finite iteration" % (stamp(), j, v))
            return True
        if j % 10 == 0 and rank == 0:
            print("%s iter = %04d, loss = %.4f" % (stamp(), j, v))
        return False

    it, stop = 0, False
    for it in range(int(args.Iteration) + 1):
        if it in eval_it_pool:                                               # distill.py:293-330
            if rank == 0:
                eng.eval_history.append((it, evaluate_synthetic_set(eng, test, args)))
                near = nearest_neighbor(sentences, eng.Y.detach(), train_text) if sentences is not None else None
                save_distilled(eng, save_path, it, near)
            if world > 1:
                torch.distributed.barrier()
        loss = eng.iteration()
        losses.push(it, loss)
        if it > 0 and read_loss(it - 1):
            stop = True
            break
    if not stop:
        read_loss(it)
    if rank == 0:
        near = nearest_neighbor(sentences, eng.Y.detach(), train_text) if sentences is not None else None
        eng.saved_to = save_distilled(eng, save_path, it, near)
    return eng


if __name__ == "__main__":
    main(parse_args())
