"""CPU oracle for the layout2img G+D hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch, functional (state-dict in, tensors out) fp32 restatement of the
reference's algorithm for the path named by BASELINE.json ``north_star``.  It exists to CHECK
the CUDA path; nothing under ``layout2img_b200/`` may import it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs use it.

Pinning: the reference ships no tests or golden vectors, so the oracle
is pinned against outputs of the *unmodified reference modules* imported from
``/root/reference`` in the build container: ``tests/golden/make_golden.py`` generated
``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` replays them (forward, gradients,
post-step buffers).  Third-party arithmetic (torch conv/linear/grid_sample/interpolate/
batch_norm, torchvision RoIAlign) is called here exactly where the reference calls it; the
versions the goldens were made with are recorded inside each golden file.

Every function cites the reference lines it follows (paths relative to /root/reference).
State is a flat ``dict[str, Tensor]`` with the reference's ``state_dict`` keys
(SURVEY.md Appendix D); in training mode the spectral-norm ``_u/_v`` vectors and the
BatchNorm running statistics in that dict are updated in place, as the reference does.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]

SN_EPS_CONV = 1e-4    # conv2d() helper: resnet_generator_app_v2.py:681-686, rcnn_discriminator_app.py:10-15
SN_EPS_DEFAULT = 1e-12  # bare nn.utils.spectral_norm(...) call sites
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
MASK_EPS = 1e-6       # norm_module.py:183-185


# --------------------------------------------------------------------------------------
# third-party state machines restated: spectral norm, batch norm
# --------------------------------------------------------------------------------------
def sn_weight(P: State, prefix: str, training: bool, eps: float) -> torch.Tensor:
    """torch.nn.utils.spectral_norm pre-forward hook (one power iteration when training).

    v <- normalize(W^T u), u <- normalize(W v) in place without grad; sigma = u^T W v with
    grad through W only; returns W / sigma.  Called once per *module call*, so shared modules
    (D.block_obj4, rcnn_discriminator_app.py:137,141) iterate twice per forward.
    """
    w = P[prefix + ".weight_orig"]
    u = P[prefix + ".weight_u"]
    v = P[prefix + ".weight_v"]
    w_mat = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            nv = torch.mv(w_mat.t(), u)
            nv = nv / nv.norm().clamp_min(eps)
            nu = torch.mv(w_mat, nv)
            nu = nu / nu.norm().clamp_min(eps)
            u.copy_(nu)
            v.copy_(nv)
            u, v = nu.clone(), nv.clone()
    sigma = torch.dot(u, torch.mv(w_mat, v))
    return w / sigma


def sn_conv(P: State, prefix: str, x, training: bool, pad: int, eps: float = SN_EPS_CONV):
    return F.conv2d(x, sn_weight(P, prefix, training, eps), P[prefix + ".bias"], 1, pad)


def sn_linear(P: State, prefix: str, x, training: bool, eps: float = SN_EPS_DEFAULT):
    return F.linear(x, sn_weight(P, prefix, training, eps), P[prefix + ".bias"])


def batch_norm(P: State, prefix: str, x, training: bool, affine: bool, count_batches: bool = False):
    """model/sync_batchnorm/batchnorm.py:48-53, single-device branch == F.batch_norm.  That
    branch calls F.batch_norm directly, so SynchronizedBatchNorm2d never advances
    ``num_batches_tracked``; only the plain nn.BatchNorm2d of the PSP stages
    (resnet_generator_app_v2.py:745) does (``count_batches=True``)."""
    w = P[prefix + ".weight"] if affine else None
    b = P[prefix + ".bias"] if affine else None
    out = F.batch_norm(x, P[prefix + ".running_mean"], P[prefix + ".running_var"], w, b,
                       training, BN_MOMENTUM, BN_EPS)
    if training and count_batches:
        P[prefix + ".num_batches_tracked"] += 1
    return out


# --------------------------------------------------------------------------------------
# a2 / a3: box-relational attention over objects
# --------------------------------------------------------------------------------------
def box_relational_embedding(bbox: torch.Tensor) -> torch.Tensor:
    """resnet_generator_app_v2.py:17-76.  bbox (b,o,4) read as (x0,y0,x1,y1) although the
    callers hand in xywh -- reference behaviour, kept.  Returns (b,o,o,64)."""
    b, o, _ = bbox.shape
    c0, c1, c2, c3 = bbox.unbind(-1)              # each (b,o)
    cx, cy = (c0 + c2) * 0.5, (c1 + c3) * 0.5
    bw, bh = (c2 - c0) + 1.0, (c3 - c1) + 1.0
    # entry [i,j]: (c_i - c_j) / size_i ; log(size_i / size_j)
    dx = torch.log(torch.clamp(torch.abs((cx[:, :, None] - cx[:, None, :]) / bw[:, :, None]), min=1e-3))
    dy = torch.log(torch.clamp(torch.abs((cy[:, :, None] - cy[:, None, :]) / bh[:, :, None]), min=1e-3))
    dw = torch.log(bw[:, :, None] / bw[:, None, :])
    dh = torch.log(bh[:, :, None] / bh[:, None, :])
    pos = torch.stack([dx, dy, dw, dh], dim=-1)  # (b,o,o,4)
    freq = 1.0 / torch.pow(torch.tensor(1000.0), torch.arange(8, dtype=torch.float32) / 8.0)
    ang = (100.0 * pos)[..., None] * freq.to(pos)   # (b,o,o,4,8)
    ang = ang.reshape(b, o, o, 32)
    return torch.cat([torch.sin(ang), torch.cos(ang)], dim=-1)


def context_attention(P: State, w: torch.Tensor, bbox: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """BoxMultiHeadedAttention(h=1, d=308, dropout=0).forward(w,w,w,bbox,mask=y)
    (resnet_generator_app_v2.py:156-214) with box_attention (:79-120)."""
    b, o, d = w.shape
    emb = box_relational_embedding(bbox.to(w))
    geo = F.relu(F.linear(emb.reshape(-1, 64), P["context.WGs.0.weight"], P["context.WGs.0.bias"]))
    geo = geo.view(b, o, o)
    q = F.linear(w, P["context.linears.0.weight"], P["context.linears.0.bias"])
    k = F.linear(w, P["context.linears.1.weight"], P["context.linears.1.bias"])
    v = F.linear(w, P["context.linears.2.weight"], P["context.linears.2.bias"])
    score = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(d)
    key_valid = (y != 0)[:, None, :].expand(b, o, o)
    score = score.masked_fill(~key_valid, -1e9)
    attn = torch.softmax(torch.log(torch.clamp(geo, min=1e-6)) + score, dim=-1)
    x = torch.matmul(attn, v)
    # :197-198 transposes (b,o,d)->(b,d,o) and then *views* the contiguous result as (b,o,d).
    # With h=1 that is not the inverse of the head split: it reinterprets the (d,o) matrix's
    # memory as (o,d), scrambling features across objects.  Reference behaviour, reproduced.
    x = x.transpose(1, 2).contiguous().view(b, o, d)
    h0 = F.layer_norm(x + w, (d,), P["context.layer_norm0.weight"], P["context.layer_norm0.bias"])
    h1 = F.linear(h0, P["context.linears.3.weight"], P["context.linears.3.bias"])
    return F.layer_norm(h1 + h0, (d,), P["context.layer_norm.weight"], P["context.layer_norm.bias"])


# --------------------------------------------------------------------------------------
# a4 / a5 / a6: mask regression, mask pasting, hard box mask
# --------------------------------------------------------------------------------------
def boxes_to_grid(boxes: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """utils/bilinear.py:161-192.  boxes (O,4) xywh -> sampling grid (O,H,W,2) in [-1,1]."""
    x0, y0, ww, hh = [boxes[:, i].view(-1, 1) for i in range(4)]
    gx = (torch.linspace(0, 1, steps=W).to(boxes).view(1, W) - x0) / ww   # (O,W)
    gy = (torch.linspace(0, 1, steps=H).to(boxes).view(1, H) - y0) / hh   # (O,H)
    O = boxes.shape[0]
    grid = torch.stack([gx[:, None, :].expand(O, H, W), gy[:, :, None].expand(O, H, W)], dim=3)
    return grid.mul(2).sub(1)


def masks_to_layout(bbox: torch.Tensor, masks: torch.Tensor, size: int) -> torch.Tensor:
    """utils/bilinear.py:137-158.  Paste (b,o,M,M) masks into (b,o,size,size) layout maps with
    grid_sample(bilinear, zeros padding, align_corners=False -- the installed torch default)."""
    b, o, M, _ = masks.shape
    if masks.shape != (b, o, M, M):
        raise AssertionError("masks must be (b, num_o, M, M)")
    grid = boxes_to_grid(bbox.reshape(b * o, 4), size, size).float()
    out = F.grid_sample(masks.float().reshape(b * o, 1, M, M), grid, mode="bilinear",
                        padding_mode="zeros", align_corners=False)
    return out.view(b, o, size, size)


def instance_norm_relu(x):
    return F.relu(F.instance_norm(x, eps=1e-5))


def mask_regress(P: State, w: torch.Tensor, bbox: torch.Tensor, training: bool) -> torch.Tensor:
    """MaskRegressNetv2.forward, model/mask_regression.py:85-102."""
    b, o, _ = bbox.shape
    pre = "mask_regress."
    x = sn_linear(P, pre + "fc", w.reshape(b * o, -1), training).view(b * o, 256, 4, 4)
    x = instance_norm_relu(sn_conv(P, pre + "conv1.0", x, training, 1, SN_EPS_DEFAULT))
    x = F.interpolate(x, size=8, mode="bilinear", align_corners=False)
    x = instance_norm_relu(sn_conv(P, pre + "conv2.0", x, training, 1, SN_EPS_DEFAULT))
    x = F.interpolate(x, size=16, mode="bilinear", align_corners=False)
    x = instance_norm_relu(sn_conv(P, pre + "conv3.0", x, training, 1, SN_EPS_DEFAULT))
    x = torch.sigmoid(sn_conv(P, pre + "conv3.3", x, training, 0, SN_EPS_DEFAULT))
    return masks_to_layout(bbox.to(w), x.view(b, o, 16, 16), 64)


def bbox_mask(bbox: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """resnet_generator_app_v2.py:697-721.  1 where 0 <= (lin - x0)/w <= 1 on both axes."""
    b, o, _ = bbox.shape
    bb = bbox.float().reshape(-1, 4)
    x0, y0, ww, hh = [bb[:, i].view(-1, 1) for i in range(4)]
    X = (torch.linspace(0, 1, steps=W).view(1, W) .to(bb) - x0) / ww
    Y = (torch.linspace(0, 1, steps=H).view(1, H).to(bb) - y0) / hh
    x_out = (X < 0) | (X > 1)
    y_out = (Y < 0) | (Y > 1)
    outside = x_out[:, None, :] | y_out[:, :, None]
    return (~outside).float().view(b, o, H, W)


# --------------------------------------------------------------------------------------
# a9: ISLA norm ; a8: generator ResBlock ; a11: mask heads ; a12: stage-mask mixing
# --------------------------------------------------------------------------------------
def isla_norm(P: State, prefix: str, x, w, mask, training: bool):
    """SpatialAdaptiveSynBatchNorm2d.forward, model/norm_module.py:163-186."""
    xh = batch_norm(P, prefix + ".batch_norm2d", x, training, affine=False)
    b, o = mask.shape[:2]
    h, wd = x.shape[2:]
    if mask.shape[2] != h or mask.shape[3] != wd:
        mask = F.interpolate(mask, size=(h, wd), mode="bilinear", align_corners=False)
    gam = sn_linear(P, prefix + ".weight_proj", w, training).view(b, o, -1)
    bet = sn_linear(P, prefix + ".bias_proj", w, training).view(b, o, -1)
    m = mask.unsqueeze(2)                                   # (b,o,1,h,w)
    denom = mask.sum(dim=1, keepdim=True) + MASK_EPS        # (b,1,h,w)
    g_pix = (m * gam[..., None, None]).sum(dim=1) / denom + 1
    b_pix = (m * bet[..., None, None]).sum(dim=1) / denom
    return g_pix * xh + b_pix


def psp_head(P: State, prefix: str, feats, training: bool, dropout_mask: Optional[torch.Tensor]):
    """PSPModule.forward, resnet_generator_app_v2.py:724-752 (sizes 1,2,3,6; out 100)."""
    h, w = feats.shape[2:]
    priors = []
    for i, s in enumerate((1, 2, 3, 6)):
        p = F.adaptive_avg_pool2d(feats, (s, s))
        p = F.conv2d(p, P[f"{prefix}.stages.{i}.1.weight"])
        p = F.relu(batch_norm(P, f"{prefix}.stages.{i}.2", p, training, affine=True, count_batches=True))
        priors.append(F.interpolate(p, size=(h, w), mode="bilinear", align_corners=True))
    priors.append(feats)
    x = F.conv2d(torch.cat(priors, 1), P[prefix + ".bottleneck.0.weight"], None, 1, 1)
    x = F.relu(batch_norm(P, prefix + ".bottleneck.1", x, training, affine=True))
    if training:
        if dropout_mask is not None:           # (b,100,1,1) keep-mask already scaled by 1/(1-p)
            x = x * dropout_mask
        else:
            x = F.dropout2d(x, 0.1, True)
    return x


def g_resblock(P: State, prefix: str, x, w, mask, training: bool, head: str,
               dropout_mask: Optional[torch.Tensor] = None):
    """ResBlock.forward (generator), resnet_generator_app_v2.py:653-678.  head in
    {"conv","psp","none"} selects the conv_mask variant (:643-651)."""
    r = F.relu(isla_norm(P, prefix + ".b1", x, w, mask, training))
    r = F.interpolate(r, scale_factor=2, mode="nearest")
    r = sn_conv(P, prefix + ".conv1", r, training, 1)
    r = F.relu(isla_norm(P, prefix + ".b2", r, w, mask, training))
    r = sn_conv(P, prefix + ".conv2", r, training, 1)
    s = sn_conv(P, prefix + ".c_sc", F.interpolate(x, scale_factor=2, mode="nearest"), training, 0)
    out = r + s
    if head == "none":
        return out, None
    cm = prefix + ".conv_mask"
    if head == "psp":
        m = psp_head(P, cm + ".0", out, training, dropout_mask)
        m = F.conv2d(m, P[cm + ".1.weight"], P[cm + ".1.bias"])
    else:
        m = F.conv2d(out, P[cm + ".0.weight"], P[cm + ".0.bias"], 1, 1)
        m = F.relu(batch_norm(P, cm + ".1", m, training, affine=True))
        m = F.conv2d(m, P[cm + ".3.weight"], P[cm + ".3.bias"])
    return out, m


def stage_mask_mix(P: State, k: int, stage_mask, bmask, hard_mask, y, size: int):
    """resnet_generator_app_v2.py:466-470 (and the three repeats below it)."""
    b, o = y.shape
    sel = torch.gather(stage_mask, 1, y.view(b, o, 1, 1).expand(b, o, size, size))
    seman = torch.sigmoid(sel) * F.interpolate(hard_mask, size=(size, size), mode="nearest")
    alpha = torch.sigmoid(P[f"alpha{k}"])[0, :, 0][y].view(b, o, 1, 1)
    soft = F.interpolate(bmask, size=(size, size), mode="bilinear", align_corners=False)
    return soft * (1 - alpha) + seman * alpha


def g_forward(P: State, z, bbox, z_im, y, training: bool, context: bool = True,
              dropout_mask: Optional[torch.Tensor] = None, taps: Optional[dict] = None):
    """ResnetGenerator128_context.forward (resnet_generator_app_v2.py:435-499); with
    ``context=False`` the plain ResnetGenerator128.forward (:333-390)."""
    b, o = z.shape[:2]
    w = torch.cat([z.reshape(b * o, -1), P["label_embedding.weight"][y].reshape(b * o, -1)], dim=1)
    if context:
        w = context_attention(P, w.view(b, o, -1), bbox, y)
    w = w.reshape(b * o, -1)
    bmask = mask_regress(P, w, bbox, training)
    if z_im is None:
        z_im = torch.randn((b, 128), device=z.device)
    hard = bbox_mask(bbox.to(z.device), 64, 64)
    x = sn_linear(P, "fc", z_im, training).view(b, -1, 4, 4)
    if taps is not None:
        taps.update(w=w, bmask=bmask, hard=hard)
    heads = ("conv", "conv", "conv", "psp", "none")
    stage = bmask
    for k in range(1, 6):
        x, sm = g_resblock(P, f"res{k}", x, w, stage, training, heads[k - 1], dropout_mask)
        if taps is not None:
            taps[f"x{k}"] = x
        if k < 5:
            stage = stage_mask_mix(P, k, sm, bmask, hard, y, x.shape[2])
            if taps is not None:
                taps[f"stage{k}"] = stage
    x = F.relu(batch_norm(P, "final.0", x, training, affine=True))
    return torch.tanh(sn_conv(P, "final.2", x, training, 1))


# --------------------------------------------------------------------------------------
# discriminator: d1 .. d6
# --------------------------------------------------------------------------------------
def d_rois(bbox: torch.Tensor, label: torch.Tensor, img_size: int):
    """CombineDiscriminator128_app.forward, rcnn_discriminator_app.py:402-417: xywh -> xyxy
    (out of place here; the reference edits a GPU-resident bbox in place), scale to pixels,
    prepend the image index, keep rows whose label != 0 in (b,o) row-major order."""
    b, o, _ = bbox.shape
    bb = bbox.float().clone()
    bb[:, :, 2] = bb[:, :, 2] + bb[:, :, 0]
    bb[:, :, 3] = bb[:, :, 3] + bb[:, :, 1]
    bb = bb * img_size
    idx = torch.arange(b, device=bb.device).view(b, 1, 1).expand(b, o, 1).float()
    rois = torch.cat([idx, bb], dim=2).view(-1, 5)
    lab = label.reshape(-1)
    keep = (lab != 0).nonzero().view(-1)
    return rois[keep], lab[keep]


def d_block(P: State, prefix: str, x, training: bool, down: bool, optimized: bool = False,
            has_sc: bool = True):
    """OptimizedBlock / ResBlock of the discriminator, rcnn_discriminator_app.py:294-344."""
    if optimized:
        r = F.relu(sn_conv(P, prefix + ".conv1", x, training, 1))
        r = sn_conv(P, prefix + ".conv2", r, training, 1)
        if down:
            r = F.avg_pool2d(r, 2)
        s = F.avg_pool2d(x, 2) if down else x
        return r + sn_conv(P, prefix + ".c_sc", s, training, 0)
    r = sn_conv(P, prefix + ".conv1", F.relu(x), training, 1)
    r = sn_conv(P, prefix + ".conv2", F.relu(r), training, 1)
    if down:
        r = F.avg_pool2d(r, 2)
    s = x
    if has_sc:
        s = sn_conv(P, prefix + ".c_sc", x, training, 0)
        if down:
            s = F.avg_pool2d(s, 2)
    return r + s


def roi_align(feat: torch.Tensor, rois: torch.Tensor, scale: float, out: int = 8):
    """torchvision.ops.RoIAlign((8,8), scale, sampling_ratio=0), aligned=False -- third-party
    (torchvision 0.26.0); call sites rcnn_discriminator_app.py:98-99,139,143."""
    from torchvision.ops import roi_align as tv_roi_align
    return tv_roi_align(feat, rois.to(feat.dtype), (out, out), scale, 0, False)   # rois are fp32; feat may be fp64 in diagnostics


def d_forward(P: State, images, bbox, label, training: bool, taps: Optional[dict] = None):
    """CombineDiscriminator128_app.forward (rcnn_discriminator_app.py:401-421) ->
    ResnetDiscriminator128_app.forward (:111-168).  State keys carry the ``obD.`` prefix."""
    rois, y = d_rois(bbox.to(images.device), label.to(images.device), images.shape[2])
    p = "obD."
    x = d_block(P, p + "block1", images, training, True, optimized=True)
    x1 = d_block(P, p + "block2", x, training, True)
    x2 = d_block(P, p + "block3", x1, training, True)
    x = d_block(P, p + "block4", x2, training, True)
    x = d_block(P, p + "block5", x, training, True)
    x = d_block(P, p + "block6", x, training, False, has_sc=False)
    out_im = sn_linear(P, p + "l7", F.relu(x).sum(dim=(2, 3)), training)

    small = ((rois[:, 3] - rois[:, 1]) < 64) & ((rois[:, 4] - rois[:, 2]) < 64)
    rois_l, rois_s = rois[~small], rois[small]
    y = torch.cat([y[~small], y[small]], dim=0)
    fs = d_block(P, p + "block_obj3", x1, training, False)
    fs = d_block(P, p + "block_obj4", fs, training, False)
    fs = roi_align(fs, rois_s, 1.0 / 4.0)
    fl = d_block(P, p + "block_obj4", x2, training, False)
    fl = roi_align(fl, rois_l, 1.0 / 8.0)
    obj = torch.cat([fl, fs], dim=0)                            # (K,512,8,8)
    if taps is not None:
        taps.update(x1=x1, x2=x2, obj=obj, rois_l=rois_l, rois_s=rois_s, y=y)

    app = F.relu(d_block(P, p + "app_conv", obj, training, False, has_sc=False))
    K, C = app.shape[:2]
    app = app.view(K, C, -1)
    gram = torch.bmm(app, app.transpose(1, 2)) / C
    ey = F.embedding(y, sn_weight(P, p + "l_y_app", training, SN_EPS_DEFAULT))
    app_all = torch.cat([gram, ey[:, None, :].expand(K, C, C)], dim=-1)
    out_app = sn_linear(P, p + "app", app_all, training).sum(1) / C

    of = F.relu(d_block(P, p + "block_obj5", obj, training, True)).sum(dim=(2, 3))
    out_obj = sn_linear(P, p + "l_obj", of, training)
    ly = F.embedding(y, sn_weight(P, p + "l_y", training, SN_EPS_DEFAULT))
    out_obj = out_obj + (ly * of).sum(dim=1, keepdim=True)
    return out_im, out_obj, out_app


# --------------------------------------------------------------------------------------
# t1: one training iteration (train_context_app_v2.py:155-189, VGG term excluded)
# --------------------------------------------------------------------------------------
LAMB_OBJ, LAMB_IMG, LAMB_APP = 1.0, 0.1, 1.0   # train_context_app_v2.py:40-46


def is_param(name: str) -> bool:
    leaf = name.rsplit(".", 1)[-1]
    return leaf not in ("weight_u", "weight_v", "running_mean", "running_var", "num_batches_tracked")


def param_names(P: State):
    return [k for k in P if is_param(k)]


def set_requires_grad(P: State, flag: bool = True):
    for k in param_names(P):
        P[k].requires_grad_(flag)


def zero_grad(P: State):
    for k in param_names(P):
        P[k].grad = None


def d_loss_fn(real_out, fake_out):
    r_im, r_obj, r_app = real_out
    f_im, f_obj, f_app = fake_out
    return (LAMB_OBJ * (F.relu(1.0 - r_obj).mean() + F.relu(1.0 + f_obj).mean())
            + LAMB_IMG * (F.relu(1.0 - r_im).mean() + F.relu(1.0 + f_im).mean())
            + LAMB_APP * (F.relu(1.0 - r_app).mean() + F.relu(1.0 + f_app).mean()))


def g_loss_fn(g_out, fake, real):
    g_im, g_obj, g_app = g_out
    return (-g_obj.mean() * LAMB_OBJ - g_im.mean() * LAMB_IMG + (fake - real).abs().mean()
            - LAMB_APP * g_app.mean())


def make_adam(P: State, lr: float):
    """One param group per tensor, betas=(0, 0.999) (train_context_app_v2.py:113-127)."""
    return torch.optim.Adam([{"params": [P[k]], "lr": lr} for k in param_names(P)], betas=(0.0, 0.999))


def train_step(PG: State, PD: State, g_opt, d_opt, real, label, bbox, z, z_im=None,
               dropout_mask=None):
    """D step then G step; returns (d_loss, g_loss, fake)."""
    zero_grad(PD)
    real_out = d_forward(PD, real, bbox, label, True)
    fake = g_forward(PG, z, bbox, z_im, label, True, dropout_mask=dropout_mask)
    fake_out = d_forward(PD, fake.detach(), bbox, label, True)
    d_loss = d_loss_fn(real_out, fake_out)
    d_loss.backward()
    d_opt.step()
    zero_grad(PG)
    g_out = d_forward(PD, fake, bbox, label, True)
    g_loss = g_loss_fn(g_out, fake, real)
    g_loss.backward()
    g_opt.step()
    return d_loss.detach(), g_loss.detach(), fake.detach()
