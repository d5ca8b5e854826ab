"""The G+D training iteration the hot path serves, restated from the reference's
train_context_app_v2.py:148-189 (VGG perceptual term optional: its weights need network access),
plus the one-process-per-GPU data-parallel plumbing (batch shard + bucketed, overlapped gradient
all-reduce per network per step).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

LAMB_OBJ, LAMB_IMG, LAMB_APP = 1.0, 0.1, 1.0       # train_context_app_v2.py:40-46


def make_optimizers(netG, netD, g_lr: float = 1e-4, d_lr: float = 1e-4, capturable: bool = False):
    """Adam(betas=(0, 0.999)) with one param group per tensor (train_context_app_v2.py:113-127; 'mapping' parameters
    would get 0.1 x lr there -- the app_v2 generator's mapping is an empty nn.Sequential), executed as one multi-tensor
    kernel per network (optim.FusedAdam -> csrc/optim.cu)."""
    from .optim import FusedAdam
    g_opt = FusedAdam([{"params": [p], "lr": g_lr * (0.1 if "mapping" in n else 1.0)} for n, p in netG.named_parameters()],
                      betas=(0.0, 0.999), capturable=capturable)
    d_opt = FusedAdam([{"params": [p], "lr": d_lr} for p in netD.parameters()], betas=(0.0, 0.999), capturable=capturable)
    return g_opt, d_opt


class _frozen:
    """Temporarily clear requires_grad on a module's parameters: in the G step the reference lets autograd
    compute D's weight gradients and then discards them at the next netD.zero_grad()
    (train_context_app_v2.py:156,178-188); skipping them changes nothing observable."""

    def __init__(self, module):
        self.params = [p for p in module.parameters() if p.requires_grad]

    def __enter__(self):
        for p in self.params:
            p.requires_grad_(False)

    def __exit__(self, *exc):
        for p in self.params:
            p.requires_grad_(True)


def _obj_mean(x, obj_scale, valid=None):
    """Mean over the (valid) objects of the GLOBAL batch.
    Single process: x.mean().  Data parallel: the reference's DataParallel gathers every replica's (K_r, 1) outputs and takes
    one mean over sum_r K_r rows (train_context_app_v2.py:159-161); with per-rank losses followed by a gradient AVERAGE over
    W ranks that is sum_local * W / sum_r K_r (obj_scale = W / sum_r K_r, a device scalar) -- identical to the plain mean
    when every rank has the same K.  Fixed-shape discriminator (static_shapes): x has b*o rows, `valid` (b*o,) flags the
    real objects; dropped rows get weight 0, so nothing flows back into them."""
    if valid is not None:
        x = x * valid.view(-1, 1).to(x.dtype)
        if obj_scale is None:
            return x.sum() / valid.sum().clamp_min(1).to(x.dtype)
        return x.sum() * obj_scale
    return x.mean() if obj_scale is None else x.sum() * obj_scale


def d_loss_fn(real_out, fake_out, obj_scale=None, valid=None):
    r_im, r_obj, r_app = real_out
    f_im, f_obj, f_app = fake_out
    om = lambda t: _obj_mean(t, obj_scale, valid)
    return (LAMB_OBJ * (om(F.relu(1.0 - r_obj)) + om(F.relu(1.0 + f_obj)))
            + LAMB_IMG * (F.relu(1.0 - r_im).mean() + F.relu(1.0 + f_im).mean())
            + LAMB_APP * (om(F.relu(1.0 - r_app)) + om(F.relu(1.0 + f_app))))


def g_loss_fn(g_out, fake, real, obj_scale=None, feat_loss=None, valid=None):
    g_im, g_obj, g_app = g_out
    loss = (-_obj_mean(g_obj, obj_scale, valid) * LAMB_OBJ - g_im.mean() * LAMB_IMG + (fake - real).abs().mean()
            - LAMB_APP * _obj_mean(g_app, obj_scale, valid))
    if feat_loss is not None:                            # train_context_app_v2.py:185-187 (VGGLoss, utils/util.py:49-94)
        loss = loss + feat_loss(fake, real).mean()
    return loss


def global_object_scale(label: torch.Tensor, group=None) -> Optional[torch.Tensor]:
    """W / (number of label != 0 objects over all ranks) as a device scalar; None on a single process."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    k = (label != 0).sum().to(torch.float32).reshape(1)
    dist.all_reduce(k, op=dist.ReduceOp.SUM, group=group)
    return (float(dist.get_world_size(group)) / k.clamp_min(1.0)).squeeze(0)


class GradBuckets:
    """Data-parallel gradient plumbing of one network (one process per GPU, replicated weights):

    * every parameter's .grad is a VIEW into one flat fp32 buffer -- autograd accumulates into it in place, the
      optimizer kernel reads it in place; there are no flatten / unflatten copies;
    * the flat buffer is laid out in the order in which the backward pass finishes the gradients and cut into buckets
      along it; a bucket's all-reduce (average) is launched on a side stream from the post-accumulate hook of its last
      parameter, so it overlaps the rest of the backward pass; `finish()` reduces whatever is left and makes the compute
      stream wait for the side stream.  The order starts as reverse registration order and is replaced, once, by the
      order OBSERVED in the first backward pass (rank 0's, broadcast) -- in the generator the mask-regression network is
      registered last but finishes last too (every stage feeds it), and a bucket waits for its slowest member.  The
      last few MB of the order form their own small bucket: it is the only all-reduce nothing can overlap;
    * parameters and buffers (spectral-norm u / v, batch-norm running statistics) are broadcast from rank 0 at
      construction, so the replicas start identical even if their initialisation was not.

    On a single process it only provides the flat gradient buffer (zero_grad = one memset).  Works with NCCL (streams,
    ReduceOp.AVG) and with gloo on CPU tensors (synchronous; used by the CPU tests)."""

    def __init__(self, module: torch.nn.Module, bucket_mb: Optional[float] = None, group=None, broadcast: bool = True,
                 tail_mb: Optional[float] = None, reorder: bool = True):
        import os
        if bucket_mb is None:
            bucket_mb = float(os.environ.get("L2I_BUCKET_MB", "48"))
        if tail_mb is None:
            tail_mb = min(float(os.environ.get("L2I_BUCKET_TAIL_MB", "4")), bucket_mb)
        self.params: List[torch.nn.Parameter] = [p for p in module.parameters() if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.dev = self.params[0].device
        self.cuda = self.dev.type == "cuda"
        self.limit = max(1, int(bucket_mb * (1 << 20) / 4))
        self.tail_limit = max(1, int(tail_mb * (1 << 20) / 4))
        self.comm_stream = torch.cuda.Stream(device=self.dev) if (self.cuda and self.world > 1) else None
        self.works = []
        if os.environ.get("L2I_BUCKET_REORDER", "1") == "0":      # A/B switch for measurements
            reorder = False
        self.reordered = not (reorder and self.world > 1)        # single process: nothing to overlap, keep the layout
        self._fired: List[int] = []
        self._observed: Optional[List[int]] = None
        self._layout(list(range(len(self.params) - 1, -1, -1)))
        if self.world > 1:
            for i, p in enumerate(self.params):
                p.register_post_accumulate_grad_hook(lambda _p, i=i: self._ready(i))
            if broadcast:
                with torch.no_grad():
                    for t in list(module.parameters()) + list(module.buffers()):
                        dist.broadcast(t, src=self._src(), group=group)

    def _src(self):
        return dist.get_global_rank(self.group, 0) if self.group is not None else 0

    def _layout(self, order: List[int]):
        """Flat buffer, .grad views and buckets for `order` (parameter indices, first-finished first)."""
        size = [(p.numel() + 3) // 4 * 4 for p in self.params]     # 16-byte aligned slices
        offs, total = [0] * len(self.params), 0
        for i in order:
            offs[i] = total
            total += size[i]
        self.order = list(order)
        self.flat = torch.zeros(total, dtype=torch.float32, device=self.dev)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(offs, self.params)]
        # the tail bucket: as many of the last-finished parameters as fit in tail_limit (at least one)
        k, acc = len(order), 0
        while k > 0 and (k == len(order) or acc + size[order[k - 1]] <= self.tail_limit):
            acc += size[order[k - 1]]
            k -= 1
        self.buckets, self.bucket_of = [], {}

        def close(members):
            if members:
                s = offs[members[0]]
                self.buckets.append({"range": (s, offs[members[-1]] + size[members[-1]]), "members": members, "pending": 0,
                                     "launched": False})
                for m in members:
                    self.bucket_of[m] = len(self.buckets) - 1

        members: List[int] = []
        for i in order[:k]:
            members.append(i)
            if sum(size[m] for m in members) >= self.limit:
                close(members)
                members = []
        close(members)
        close(list(order[k:]))
        self._attach()
        self._reset()

    def _attach(self):
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                p.grad = v

    def _reset(self):
        for b in self.buckets:
            b["pending"], b["launched"] = len(b["members"]), False
        self.works = []
        self._fired = []

    def _maybe_reorder(self):
        """Once, at the first zero_grad after a complete backward pass (the gradients are about to be cleared, so nothing
        has to be carried over): lay the buffer out in the observed completion order.  Every rank uses rank 0's order."""
        if self.reordered or self._observed is None or (self.cuda and torch.cuda.is_current_stream_capturing()):
            return
        seen = set(self._observed)
        order = self._observed + [i for i in self.order if i not in seen]       # never-fired parameters keep their place at the end
        t = torch.tensor(order, dtype=torch.int64, device=self.dev)
        dist.broadcast(t, src=self._src(), group=self.group)
        order = [int(i) for i in t.tolist()]
        if sorted(order) != list(range(len(self.params))):
            raise RuntimeError("GradBuckets: rank 0 observed an inconsistent gradient order")
        self.reordered = True
        if order != self.order:
            if self.comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
            self._layout(order)

    def zero_grad(self):
        """Replaces module.zero_grad(): one memset of the flat buffer; the .grad views stay attached."""
        self._maybe_reorder()
        self.flat.zero_()
        self._attach()

    def _launch(self, b):
        b["launched"] = True
        s, e = b["range"]
        chunk = self.flat[s:e]
        if self.comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
            chunk.div_(self.world)

    def _ready(self, i: int):
        self._fired.append(i)
        b = self.buckets[self.bucket_of[i]]
        b["pending"] -= 1
        if b["pending"] == 0 and not b["launched"]:
            self._launch(b)

    def finish(self):
        """Call after backward(): reduce the buckets whose hooks did not all fire (parameters without a gradient in this
        pass count as zeros), then order the compute stream after the side stream."""
        if self.world > 1:
            for b in self.buckets:
                if not b["launched"]:
                    self._launch(b)
            if self.comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
            if not self.reordered and self._observed is None and self._fired:
                self._observed = list(dict.fromkeys(self._fired))
        self._reset()

    __call__ = finish


# Backwards-compatible name (round 1 exposed the post-backward flat all-reduce under it).
GradAllReducer = GradBuckets


def setup_data_parallel(netG, netD, sync_bn: bool = True, group=None):
    """One call for a rank of a data-parallel job: gradient buckets for both networks (with the rank-0 broadcast) and,
    by default, batch-norm statistics over the GLOBAL batch -- the reference's multi-GPU semantics
    (SynchronizedBatchNorm2d under DataParallelWithCallback, sync_batchnorm/batchnorm.py:90-125).  sync_bn=False keeps
    per-rank statistics (= the reference run with the per-GPU batch on one GPU)."""
    from . import ops
    ops.set_sync_bn(bool(sync_bn), group)
    return GradBuckets(netG, group=group), GradBuckets(netD, group=group)


def train_step(netG, netD, g_opt, d_opt, real, label, bbox, z, z_im=None, sync_g=None, sync_d=None,
               record=None, feat_loss=None):
    """One D step then one G step; returns (d_loss, g_loss, fake) as detached tensors (no host sync).
    sync_g / sync_d: GradBuckets of the two networks (data parallel, or single process for the flat gradient buffer).
    `record(tag)` is an optional callback used by the parity tests to snapshot gradients; feat_loss an optional
    perceptual loss module (VGGLoss)."""
    lab3 = label.unsqueeze(-1) if label.dim() == 2 else label
    obj_scale = global_object_scale(label, sync_d.group if sync_d is not None else None) \
        if (sync_d is not None and sync_d.world > 1) else None
    # ---- D step (:155-174)
    if sync_d is not None:
        sync_d.zero_grad()
    else:
        netD.zero_grad()
    real_out = netD(real, bbox, lab3)
    valid = getattr(netD, "valid_mask", None)        # fixed-shape discriminator: flags of the real objects (else None)
    fake = netG(z, bbox, z_im, y=label.view(label.shape[0], -1))
    fake_out = netD(fake.detach(), bbox, lab3)
    d_loss = d_loss_fn(real_out, fake_out, obj_scale, valid)
    d_loss.backward()
    if sync_d is not None:
        sync_d.finish()
    if record is not None:
        record("d")
    d_opt.step()
    # ---- G step (:177-189); D's gradients produced here are discarded by the next zero_grad
    if sync_g is not None:
        sync_g.zero_grad()
    else:
        netG.zero_grad()
    with _frozen(netD):
        g_out = netD(fake, bbox, lab3)
        g_loss = g_loss_fn(g_out, fake, real, obj_scale, feat_loss, valid)
        g_loss.backward()
    if sync_g is not None:
        sync_g.finish()
    if record is not None:
        record("g")
    g_opt.step()
    return d_loss.detach(), g_loss.detach(), fake.detach()


class GraphedTrainStep:
    """The whole training iteration (D step + G step, both optimizer updates) recorded ONCE as a CUDA graph and replayed
    from static input buffers: one graph launch per step instead of ~1 500 kernel launches issued from Python, and no
    host synchronisation anywhere.  Requirements, all arranged here:

      * the discriminator runs in its fixed-shape form (`static_shapes`: device-side ROI compaction, dropped objects as
        zero-feature rows that the masked losses ignore) -- the eager form's output size depends on the labels;
      * gradients live at fixed addresses (GradBuckets views) and the optimizers are `FusedAdam(capturable=True)` (step
        count on the device, static tensor table).

    Numerically the replay IS the eager fixed-shape step (same kernels, same order).  `__call__` copies the batch into the
    static buffers (device or pinned-host sources) and replays; it returns the static (d_loss, g_loss, fake) tensors,
    valid until the next call."""

    def __init__(self, netG, netD, g_opt, d_opt, real, label, bbox, z, z_im, warmup: int = 3, feat_loss=None):
        if not (getattr(g_opt, "capturable", False) and getattr(d_opt, "capturable", False)):
            raise ValueError("GraphedTrainStep needs FusedAdam(..., capturable=True) optimizers (make_optimizers(capturable=True))")
        dev = real.device
        if dev.type != "cuda":
            raise RuntimeError("layout2img_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        netD.static_shapes = True
        self.netG, self.netD, self.g_opt, self.d_opt, self.feat_loss = netG, netD, g_opt, d_opt, feat_loss
        self.sync_g, self.sync_d = GradBuckets(netG), GradBuckets(netD)
        self.static = {"real": real.clone(), "label": label.clone(), "bbox": bbox.to(dev).float().clone(), "z": z.clone(),
                       "z_im": z_im.clone()}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        # eager steps before the capture: they build every lazily created table / workspace, and (data parallel) the
        # second one re-cuts the gradient buckets in the completion order the first one observed
        warmup = max(2 if self.sync_d.world > 1 else 1, warmup)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from ._lib import lib
        self.graph = torch.cuda.CUDAGraph()
        lib().l2i_launch_count(1)
        with torch.cuda.graph(self.graph):
            self.out = self._step()
        self.kernels_per_replay = lib().l2i_launch_count(1)      # libl2i kernel nodes recorded in the graph
        self.warmup_steps = warmup                # eager steps already applied to the networks (capture itself executes nothing)

    def _step(self):
        s = self.static
        return train_step(self.netG, self.netD, self.g_opt, self.d_opt, s["real"], s["label"], s["bbox"], s["z"], s["z_im"],
                          sync_g=self.sync_g, sync_d=self.sync_d, feat_loss=self.feat_loss)

    def __call__(self, real, label, bbox, z, z_im):
        s = self.static
        for k, v in (("real", real), ("label", label), ("bbox", bbox), ("z", z), ("z_im", z_im)):
            if v is not s[k]:
                s[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.out
