"""Per-operator parity on the B200: every libl2i.so kernel against the oracle's restatement of the
same reference lines (CPU, fp32/fp64) or torch autograd of that restatement.  All calls go through
the C ABI (layout2img_b200.ops -> ctypes -> libl2i.so)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_case
from layout2img_b200.synth import synthetic_layout
from oracle import l2i_oracle as O

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-4


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def close(got, want, rtol=RTOL, atol=ATOL, what=""):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} outside tol, max err {err.max().item():.3e} (ref max {want.abs().max().item():.3e})"


CONV_SHAPES = [  # N, Cin, Cout, H, k   (small instances of every channel/tiling class of SURVEY.md App. A)
    (2, 64, 64, 16, 3), (2, 64, 128, 16, 1), (3, 128, 128, 8, 3), (5, 256, 256, 4, 3), (2, 3, 64, 32, 3),
    (2, 64, 3, 32, 3), (2, 256, 100, 16, 3), (1, 528, 100, 16, 3), (2, 100, 184, 8, 1), (7, 512, 1024, 8, 3),
    (2, 256, 1, 16, 1), (2, 128, 64, 64, 3),
]


@pytest.mark.parametrize("N,Cin,Cout,H,k", CONV_SHAPES)
def test_conv_fwd_dgrad_wgrad(dev, N, Cin, Cout, H, k):
    from layout2img_b200 import functional as L
    g = torch.Generator().manual_seed(N * 1000 + Cin + Cout + H)
    x = torch.randn(N, Cin, H, H, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g)
    dy = torch.randn(N, Cout, H, H, generator=g)
    xr, wr, br = x.double().requires_grad_(), w.double().requires_grad_(), b.double().requires_grad_()
    ref = F.conv2d(xr, wr, br, padding=k // 2)
    ref.backward(dy.double())
    xg = nhwc(x).to(dev).requires_grad_()
    wg = w.to(dev).requires_grad_()
    bg = b.to(dev).requires_grad_()
    out = L.conv2d(xg, wg, bg)
    out.backward(nhwc(dy).to(dev))
    scale = ref.abs().max().item()
    close(out.permute(0, 3, 1, 2), ref, 1e-3, 1e-4 * max(scale, 1), "fwd")
    close(xg.grad.permute(0, 3, 1, 2), xr.grad, 1e-3, 1e-4 * max(xr.grad.abs().max().item(), 1), "dgrad")
    close(wg.grad, wr.grad, 1e-3, 2e-4 * max(wr.grad.abs().max().item(), 1), "wgrad")
    close(bg.grad, br.grad, 1e-3, 1e-4 * max(br.grad.abs().max().item(), 1), "dbias")


def test_conv_prologue_epilogue_fusions(dev):
    """relu-in, nearest-x2-in, residual (same and half resolution) == the unfused composition."""
    from layout2img_b200 import functional as L
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 64, 8, 8, generator=g)
    w = torch.randn(128, 64, 3, 3, generator=g) / 24
    r_lo = torch.randn(2, 128, 8, 8, generator=g)
    xr, wr, rr = x.double().requires_grad_(), w.double().requires_grad_(), r_lo.double().requires_grad_()
    ref = F.conv2d(F.interpolate(F.relu(xr), scale_factor=2, mode="nearest"), wr, None, padding=1) + \
        F.interpolate(rr, scale_factor=2, mode="nearest")
    dy = torch.randn(ref.shape, generator=g)
    ref.backward(dy.double())
    xg, wg, rg = nhwc(x).to(dev).requires_grad_(), w.to(dev).requires_grad_(), nhwc(r_lo).to(dev).requires_grad_()
    out = L.conv2d(xg, wg, None, rg, relu_in=True, up2_in=True, res_up2=True)
    out.backward(nhwc(dy).to(dev))
    close(out.permute(0, 3, 1, 2), ref, what="fwd")
    close(xg.grad.permute(0, 3, 1, 2), xr.grad, what="dx")
    close(rg.grad.permute(0, 3, 1, 2), rr.grad, what="dres")
    close(wg.grad, wr.grad, 1e-3, 1e-3, what="dw")


@pytest.mark.parametrize("N,Cin,Cout,H,pool,taps", [(3, 64, 128, 16, 1, 9), (2, 128, 64, 8, 2, 9), (9, 72, 100, 4, 1, 1),
                                                   (2, 64, 3, 32, 2, 9), (150, 64, 128, 16, 1, 9)])
def test_conv_epilogue_mask_pool_residual_pair(dev, N, Cin, Cout, H, pool, taps):
    """The fused epilogue of l2i_conv2d_fwd: ReLU-derivative mask from a saved bf16 activation, 2x2
    average/sum pooling, scaled residual at the stored resolution, (ReLU'd) pair output -- against the
    unfused fp64 composition.  N=150 gives more tiles than SMs (persistent loop, both TMEM buffers)."""
    from layout2img_b200 import ops
    g = torch.Generator().manual_seed(N + Cin + Cout + H + pool)
    k = 3 if taps == 9 else 1
    x = torch.randn(N, Cin, H, H, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * taps) ** 0.5
    b = torch.randn(Cout, generator=g)
    act = torch.randn(N, Cout, H, H, generator=g)                    # the saved activation whose sign masks
    res = torch.randn(N, Cout, H // 2, H // 2, generator=g)
    v = F.conv2d(x.double(), w.double(), b.double(), padding=k // 2) * (act.double() > 0)
    v = F.avg_pool2d(v, 2) * (1.0 if pool == 1 else 4.0)
    ref = v + 0.25 * res.double()
    xp = ops.act_split(nhwc(x).to(dev))
    wp = ops.conv_weight_prep(w.to(dev), need_dgrad=False)
    mask = ops.act_split(nhwc(act).to(dev), relu=True)               # mask_hi = hi half of relu(act)
    out, pair = ops.conv2d_fwd(xp, wp.f_hi, wp.f_lo, Cout, taps, bias=b.to(dev), residual=nhwc(res).to(dev),
                               res_scale=0.25, mask_hi=mask.hi, pool=pool, want_f32=True, want_pair=True, relu_pair=True)
    scale = max(ref.abs().max().item(), 1.0)
    close(out.permute(0, 3, 1, 2), ref, 1e-3, 1e-4 * scale, "pooled masked conv")
    got_pair = (pair.hi.float() + pair.lo.float())[..., :Cout].permute(0, 3, 1, 2)
    close(got_pair, F.relu(ref), 1e-3, 1e-4 * scale, "relu pair output")
    assert pair.hi.shape[-1] % 8 == 0 and float(pair.hi[..., Cout:].abs().sum()) == 0.0


@pytest.mark.parametrize("name", ["C", "Cpad", "V"])
def test_bbox_mask_bit_exact_against_reference_golden(dev, name):
    from layout2img_b200 import ops
    z, meta = load_case(name)
    data = synthetic_layout(meta["batch"], meta["num_obj"], meta["num_classes"], seed=meta["seed"], n_pad=meta["n_pad"])
    for size in (64, 128):
        got = ops.bbox_mask(data["bbox"].to(dev), size, size).cpu()
        assert torch.equal(got, O.bbox_mask(data["bbox"], size, size)), "differs from oracle"
        assert np.array_equal(np.packbits(got.numpy().astype(np.uint8)), z[f"op.bbox_mask{size}"]), "differs from reference golden"


@pytest.mark.parametrize("name", ["C", "Cpad", "V"])
def test_masks_to_layout_fwd_bwd(dev, name):
    from layout2img_b200 import functional as L
    z, meta = load_case(name)
    data = synthetic_layout(meta["batch"], meta["num_obj"], meta["num_classes"], seed=meta["seed"], n_pad=meta["n_pad"])
    gm = torch.Generator().manual_seed(meta["seed"] + 100)
    masks = torch.rand(meta["batch"], meta["num_obj"], 16, 16, generator=gm)
    mr = masks.clone().requires_grad_()
    ref = O.masks_to_layout(data["bbox"], mr, 64)
    dy = torch.randn(ref.shape, generator=gm)
    ref.backward(dy)
    mg = masks.to(dev).requires_grad_()
    out = L.masks_to_layout(mg, data["bbox"].to(dev), 64)
    out.backward(dy.to(dev))
    np.testing.assert_allclose(out.detach().cpu().numpy(), z["op.masks_to_layout"], rtol=1e-5, atol=1e-5)
    close(out, ref, 1e-5, 1e-5, "fwd")
    close(mg.grad, mr.grad, 1e-4, 1e-4, "bwd")


@pytest.mark.parametrize("hi,h", [(64, 4), (64, 8), (64, 64), (8, 16), (32, 64), (64, 128)])
@pytest.mark.parametrize("pm", [True, False])
def test_mask_resize_fwd_bwd(dev, hi, h, pm):
    from layout2img_b200 import functional as L
    g = torch.Generator().manual_seed(hi + h)
    m = torch.rand(2, 5, hi, hi, generator=g)
    mr = m.clone().requires_grad_()
    ref = F.interpolate(mr, size=(h, h), mode="bilinear", align_corners=False) if hi != h else mr * 1.0
    dy = torch.randn(ref.shape, generator=g)
    ref.backward(dy)
    mg = m.to(dev).requires_grad_()
    out = L.mask_resize(mg, h, h, pm)
    dyg = dy.to(dev)
    out.backward(dyg.permute(0, 2, 3, 1).contiguous() if pm else dyg)
    got = out.permute(0, 3, 1, 2) if pm else out
    close(got, ref, 1e-5, 1e-6, "fwd")
    close(mg.grad, mr.grad, 1e-4, 1e-5, "bwd")


def _isla_ref(x, mask, gamma, beta, rm, rv, training):
    xh = F.batch_norm(x, rm, rv, None, None, training, 0.1, 1e-5)
    m = mask.unsqueeze(2)
    den = mask.sum(1, keepdim=True) + 1e-6
    g = (m * gamma[..., None, None]).sum(1) / den + 1
    b = (m * beta[..., None, None]).sum(1) / den
    return g * xh + b


@pytest.mark.parametrize("B,O,C,H,up,training", [(2, 4, 64, 16, True, True), (3, 8, 256, 8, False, True),
                                                 (2, 16, 128, 4, True, True), (2, 4, 1024, 4, True, False),
                                                 (2, 31, 64, 8, False, True)])
def test_isla_norm_relu_conv_fwd_bwd(dev, B, O, C, H, up, training):
    """Fused ISLA + ReLU (+ nearest x2) + 3x3 conv against the unfused fp64 composition
    (norm_module.py:163-186 + ResBlock.residual), forward, every gradient and the running stats."""
    from layout2img_b200 import functional as L
    g = torch.Generator().manual_seed(B * 100 + O + C)
    x = torch.randn(B, C, H, H, generator=g) * 1.5 + 0.3
    mask = torch.rand(B, O, H, H, generator=g) * (torch.rand(B, O, H, H, generator=g) > 0.4)
    gamma, beta = torch.randn(B, O, C, generator=g) * 0.3, torch.randn(B, O, C, generator=g) * 0.3
    w = torch.randn(64, C, 3, 3, generator=g) / (3 * C ** 0.5)
    bias = torch.randn(64, generator=g) * 0.1
    rm, rv = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
    dbl = [t.double().requires_grad_() for t in (x, mask, gamma, beta, w, bias)]
    rm_r, rv_r = rm.double().clone(), rv.double().clone()
    a = F.relu(_isla_ref(dbl[0], dbl[1], dbl[2], dbl[3], rm_r, rv_r, training))
    if up:
        a = F.interpolate(a, scale_factor=2, mode="nearest")
    ref = F.conv2d(a, dbl[4], dbl[5], padding=1)
    dy = torch.randn(ref.shape, generator=g)
    ref.backward(dy.double())
    xg = nhwc(x).to(dev).requires_grad_()
    mg = mask.permute(0, 2, 3, 1).contiguous().to(dev).requires_grad_()
    gg, bg, wg, biasg = [t.to(dev).requires_grad_() for t in (gamma, beta, w, bias)]
    rmg, rvg = rm.to(dev), rv.to(dev)
    out = L.norm_conv(xg, wg, biasg, rmg, rvg, training, mask_pm=mg, gamma=gg, beta=bg, up2=up)
    out.backward(nhwc(dy).to(dev))
    close(out.permute(0, 3, 1, 2), ref, what="fwd")
    close(rmg, rm_r, 1e-4, 1e-5, "running_mean")
    close(rvg, rv_r, 1e-4, 1e-5, "running_var")
    close(xg.grad.permute(0, 3, 1, 2), dbl[0].grad, 1e-3, 2e-4, "dx")
    dm = dbl[1].grad
    close(mg.grad.permute(0, 3, 1, 2), dm, 2e-3, 1e-4 + 1e-5 * dm.abs().max().item(), "dmask")
    close(gg.grad, dbl[2].grad, 1e-3, 1e-3, "dgamma")
    close(bg.grad, dbl[3].grad, 1e-3, 1e-3, "dbeta")
    close(wg.grad, dbl[4].grad, 1e-3, 1e-3, "dw")
    close(biasg.grad, dbl[5].grad, 1e-3, 1e-3, "dbias")


@pytest.mark.parametrize("training", [True, False])
def test_affine_bn_relu_conv(dev, training):
    """BN(affine) -> ReLU -> conv (final / mask heads, resnet_generator_app_v2.py:416-419,645-651)."""
    from layout2img_b200 import functional as L
    g = torch.Generator().manual_seed(9)
    C = 100
    x = torch.randn(3, C, 8, 8, generator=g) + 0.2
    aw, ab = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.2
    w = torch.randn(184, C, 1, 1, generator=g) / 10
    bias = torch.randn(184, generator=g) * 0.1
    rm, rv = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
    dbl = [t.double().requires_grad_() for t in (x, aw, ab, w, bias)]
    rm_r, rv_r = rm.double().clone(), rv.double().clone()
    ref = F.conv2d(F.relu(F.batch_norm(dbl[0], rm_r, rv_r, dbl[1], dbl[2], training, 0.1, 1e-5)), dbl[3], dbl[4])
    dy = torch.randn(ref.shape, generator=g)
    ref.backward(dy.double())
    xg = nhwc(x).to(dev).requires_grad_()
    awg, abg, wg, biasg = [t.to(dev).requires_grad_() for t in (aw, ab, w, bias)]
    out = L.norm_conv(xg, wg, biasg, rm.to(dev), rv.to(dev), training, aff_w=awg, aff_b=abg)
    out.backward(nhwc(dy).to(dev))
    close(out.permute(0, 3, 1, 2), ref, what="fwd")
    close(xg.grad.permute(0, 3, 1, 2), dbl[0].grad, 1e-3, 2e-4, "dx")
    close(awg.grad, dbl[1].grad, 1e-3, 1e-3, "daffw")
    close(abg.grad, dbl[2].grad, 1e-3, 1e-3, "daffb")
    close(wg.grad, dbl[3].grad, 1e-3, 1e-3, "dw")


@pytest.mark.parametrize("name", ["C", "Cpad", "V"])
def test_stage_mask_mix_fwd_bwd(dev, name):
    from layout2img_b200 import functional as L
    z, meta = load_case(name)
    data = synthetic_layout(meta["batch"], meta["num_obj"], meta["num_classes"], seed=meta["seed"], n_pad=meta["n_pad"])
    b, o = meta["batch"], meta["num_obj"]
    g = torch.Generator().manual_seed(3)
    hard = O.bbox_mask(data["bbox"], 64, 64)
    bmask = torch.rand(b, o, 64, 64, generator=g)
    alpha = torch.randn(1, 184, 1, generator=g)
    for size in (8, 64):
        stage = torch.randn(b, 184, size, size, generator=g)
        sr, ar, br = stage.clone().requires_grad_(), alpha.clone().requires_grad_(), bmask.clone().requires_grad_()
        ref = O.stage_mask_mix({"alpha1": ar}, 1, sr, br, hard, data["label"], size)
        dy = torch.randn(ref.shape, generator=g)
        ref.backward(dy)
        sg = nhwc(stage).to(dev).requires_grad_()
        ag, bg = alpha.to(dev).requires_grad_(), bmask.to(dev).requires_grad_()
        out = L.stage_mix(sg, ag, bg, data["label"].to(dev), hard.to(dev))
        out.backward(dy.to(dev))
        close(out, ref, 1e-4, 1e-5, "fwd")
        close(sg.grad.permute(0, 3, 1, 2), sr.grad, 1e-3, 1e-5, "dstage")
        close(ag.grad, ar.grad, 1e-3, 1e-4, "dalpha")
        close(bg.grad, br.grad, 1e-3, 1e-5, "dbmask")


@pytest.mark.parametrize("name", ["C", "Cpad", "V"])
def test_box_attention_fwd_bwd(dev, name):
    """Kernel vs the oracle's restatement of :17-120,172-192 (embedding, gate, masked softmax, PV)."""
    from layout2img_b200 import functional as L
    z, meta = load_case(name)
    data = synthetic_layout(meta["batch"], meta["num_obj"], meta["num_classes"], seed=meta["seed"], n_pad=meta["n_pad"])
    b, o, d = meta["batch"], meta["num_obj"], 308
    g = torch.Generator().manual_seed(21)
    q, k, v = [torch.randn(b, o, d, generator=g) for _ in range(3)]
    wgw, wgb = torch.randn(1, 64, generator=g) * 0.3, torch.randn(1, generator=g) * 0.1 + 0.5
    leaves = [t.clone().requires_grad_() for t in (q, k, v, wgw, wgb)]
    emb = O.box_relational_embedding(data["bbox"])
    np.testing.assert_allclose(emb.numpy(), z["op.box_rel_emb"], rtol=1e-5, atol=1e-5)
    geo = F.relu(F.linear(emb.reshape(-1, 64), leaves[3], leaves[4])).view(b, o, o)
    score = torch.matmul(leaves[0], leaves[1].transpose(-2, -1)) / d ** 0.5
    score = score.masked_fill(~(data["label"] != 0)[:, None, :].expand(b, o, o), -1e9)
    ref = torch.matmul(torch.softmax(torch.log(torch.clamp(geo, min=1e-6)) + score, dim=-1), leaves[2])
    dy = torch.randn(ref.shape, generator=g)
    ref.backward(dy)
    gl = [t.to(dev).requires_grad_() for t in (q, k, v, wgw, wgb)]
    out = L.box_attention(gl[0], gl[1], gl[2], data["bbox"].to(dev), data["label"].to(dev), gl[3], gl[4])
    out.backward(dy.to(dev))
    close(out, ref, what="fwd")
    for nm, a, r in zip(("dq", "dk", "dv", "dwg", "dbg"), gl, leaves):
        close(a.grad, r.grad, 1e-3, 2e-4, nm)


@pytest.mark.parametrize("H,scale", [(32, 0.25), (16, 0.125)])
def test_roi_align_fwd_bwd(dev, H, scale):
    """vs torchvision.ops.roi_align (the third-party op the reference calls, rcnn_discriminator_app.py:98-99)."""
    from torchvision.ops import roi_align as tv_roi_align
    from layout2img_b200 import functional as L
    data = synthetic_layout(4, 8, 184, seed=4)
    rois, _ = O.d_rois(data["bbox"], data["label"], 128)
    # include degenerate / border boxes
    extra = torch.tensor([[0, 0, 0, 128, 128], [1, 100, 100, 100.5, 100.2], [2, 120, 3, 128, 60], [3, 0, 0, 1, 1]], dtype=torch.float32)
    rois = torch.cat([rois, extra])
    g = torch.Generator().manual_seed(8)
    feat = torch.randn(4, 64, H, H, generator=g)
    fr = feat.clone().requires_grad_()
    ref = tv_roi_align(fr, rois, (8, 8), scale, 0, False)
    dy = torch.randn(ref.shape, generator=g)
    ref.backward(dy)
    fg = nhwc(feat).to(dev).requires_grad_()
    out = L.roi_align(fg, rois.to(dev), scale)
    out.backward(nhwc(dy).to(dev))
    close(out.permute(0, 3, 1, 2), ref, 1e-4, 1e-5, "fwd")
    close(fg.grad.permute(0, 3, 1, 2), fr.grad, 1e-4, 1e-4, "bwd")


def test_roi_align_empty(dev):
    from layout2img_b200 import functional as L
    fg = torch.randn(2, 16, 16, 64, device=dev, requires_grad=True)
    out = L.roi_align(fg, torch.zeros((0, 5), device=dev), 0.125)
    assert out.shape == (0, 8, 8, 64)
    out.sum().backward()
    assert torch.count_nonzero(fg.grad) == 0


def test_avgpool(dev):
    from layout2img_b200 import functional as L
    x = torch.randn(2, 64, 16, 16)
    xr = x.clone().requires_grad_()
    ref = F.avg_pool2d(xr, 2)
    dy = torch.randn(ref.shape)
    ref.backward(dy)
    xg = nhwc(x).to(dev).requires_grad_()
    out = L.avgpool2(xg)
    out.backward(nhwc(dy).to(dev))
    close(out.permute(0, 3, 1, 2), ref, 1e-6, 1e-6)
    close(xg.grad.permute(0, 3, 1, 2), xr.grad, 1e-6, 1e-6)


def test_bad_arguments_raise(dev):
    from layout2img_b200 import ops
    x = torch.zeros(1, 6, 6, 8, device=dev)      # H, W not powers of two
    wp = ops.conv_weight_prep(torch.zeros(8, 8, 3, 3, device=dev))
    with pytest.raises(ValueError):
        ops.conv2d_fwd(ops.act_split(x), wp.f_hi, wp.f_lo, 8, 9)
    with pytest.raises(RuntimeError):
        ops.act_split(torch.zeros(1, 4, 4, 8))   # CPU tensor: no fallback


def test_fused_adam_matches_torch_adam(dev):
    """optim.FusedAdam (one launch for all tensors) vs torch.optim.Adam(betas=(0, 0.999)) with one group per
    tensor, as the reference builds it (train_context_app_v2.py:113-127): 4 steps, odd sizes, a tensor
    without gradient."""
    from layout2img_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(3)
    shapes = [(1,), (3,), (100, 7), (184, 100, 1, 1), (50001,), (64, 64, 3, 3), (5,)]
    ps_a = [torch.randn(s, generator=g).to(dev).requires_grad_() for s in shapes]
    ps_b = [p.detach().clone().requires_grad_() for p in ps_a]
    lrs = [1e-4 * (1 + i) for i in range(len(shapes))]
    a = FusedAdam([{"params": [p], "lr": lr} for p, lr in zip(ps_a, lrs)], betas=(0.0, 0.999))
    b = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(ps_b, lrs)], betas=(0.0, 0.999))
    for step in range(4):
        for i, (pa, pb) in enumerate(zip(ps_a, ps_b)):
            if i == len(shapes) - 1:
                continue                        # never receives a gradient
            gr = (torch.randn(pa.shape, generator=g) * 10 ** float(step - 2)).to(dev)
            pa.grad, pb.grad = gr.clone(), gr.clone()
        a.step(); b.step()
    for pa, pb in zip(ps_a, ps_b):
        close(pa, pb, rtol=1e-6, atol=1e-7, what="adam param")
    for pa, pb in zip(ps_a[:-1], ps_b[:-1]):
        close(a.state[pa]["exp_avg_sq"], b.state[pb]["exp_avg_sq"], rtol=1e-6, atol=1e-12, what="exp_avg_sq")


def test_fused_adam_per_tensor_step_rebuild_and_state_dict(dev):
    """(i) a parameter whose gradient is missing in some steps keeps its OWN step count (torch's per-parameter bias
    correction); (ii) un-freezing a parameter later rebuilds the flat buffers WITHOUT wiping the other tensors' moments;
    (iii) state_dict()/load_state_dict() interoperate with torch.optim.Adam in both directions; (iv) the raw-pointer
    update bumps autograd's version counter."""
    from layout2img_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(5)
    shapes = [(33,), (17, 9), (4, 4, 3, 3)]
    mk = lambda: [torch.randn(s, generator=torch.Generator().manual_seed(7 + i)).to(dev).requires_grad_() for i, s in enumerate(shapes)]
    ps_a, ps_b = mk(), mk()
    ps_a[2].requires_grad_(False); ps_b[2].requires_grad_(False)          # frozen at first
    a = FusedAdam([{"params": [p], "lr": 1e-3} for p in ps_a], betas=(0.0, 0.999))
    b = torch.optim.Adam([{"params": [p], "lr": 1e-3} for p in ps_b], betas=(0.0, 0.999))
    v0 = ps_a[0]._version
    for step in range(6):
        if step == 3:
            ps_a[2].requires_grad_(True); ps_b[2].requires_grad_(True)  # joins later: plan rebuild
        for i, (pa, pb) in enumerate(zip(ps_a, ps_b)):
            pa.grad = pb.grad = None
            if not pa.requires_grad or (i == 1 and step in (1, 2)):       # tensor 1 skips two steps
                continue
            gr = torch.randn(pa.shape, generator=g).to(dev)
            pa.grad, pb.grad = gr.clone(), gr.clone()
        a.step(); b.step()
    assert ps_a[0]._version > v0
    for pa, pb in zip(ps_a, ps_b):
        close(pa, pb, rtol=1e-6, atol=1e-7, what="adam param after rebuild / skipped steps")
        assert a.state[pa]["step"] == int(b.state[pb]["step"])
    # torch -> ours and ours -> torch
    ps_c, ps_d = [p.detach().clone().requires_grad_() for p in ps_a], [p.detach().clone().requires_grad_() for p in ps_a]
    c = FusedAdam([{"params": [p], "lr": 1e-3} for p in ps_c], betas=(0.0, 0.999))
    d = torch.optim.Adam([{"params": [p], "lr": 1e-3} for p in ps_d], betas=(0.0, 0.999))
    c.load_state_dict(b.state_dict())
    d.load_state_dict(a.state_dict())
    for pc, pd in zip(ps_c, ps_d):
        gr = torch.randn(pc.shape, generator=g).to(dev)
        pc.grad, pd.grad = gr.clone(), gr.clone()
    c.step(); d.step()
    for pc, pd in zip(ps_c, ps_d):
        close(pc, pd, rtol=1e-6, atol=1e-7, what="adam param after state_dict round trip")


@pytest.mark.parametrize("cin,cout,H,down,optimized", [(3, 64, 32, True, True), (64, 128, 16, True, False),
                                                      (128, 256, 8, False, False), (72, 72, 8, False, False),
                                                      (512, 1024, 8, True, False)])
def test_d_block_fused_matches_composition(dev, cin, cout, H, down, optimized):
    """functional.d_block (one autograd node: operand prep + 3 convs with fused ReLU / pooling / shortcut) vs
    the reference composition of rcnn_discriminator_app.py:294-344 in fp64, forward and every gradient."""
    from layout2img_b200 import functional as L
    N = 3
    has_sc = down or cin != cout
    g = torch.Generator().manual_seed(cin + cout + H)
    x = torch.randn(N, cin, H, H, generator=g)
    w1 = torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5
    w2 = torch.randn(cout, cout, 3, 3, generator=g) / (9 * cout) ** 0.5
    wsc = torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5
    b1, b2, bsc = (torch.randn(cout, generator=g) for _ in range(3))
    leaves = [t.double().requires_grad_() for t in (x, w1, b1, w2, b2, wsc, bsc)]
    xr, w1r, b1r, w2r, b2r, wscr, bscr = leaves
    # The inner ReLU sits on a K = 9*cin reduction: a pre-activation within rounding distance of 0 may land on
    # either side of the kink in two correct implementations.  Give the fp64 composition the side the kernels
    # took (the unfused conv produces bit-identical accumulators), so the comparison tests the arithmetic.
    with torch.no_grad():
        h1 = L.conv2d(nhwc(x).to(dev), w1.to(dev), b1.to(dev), relu_in=not optimized)
        m1 = (h1 > 0).permute(0, 3, 1, 2).double().cpu()
    if optimized:
        r = F.conv2d(F.conv2d(xr, w1r, b1r, padding=1) * m1, w2r, b2r, padding=1)
        ref = F.avg_pool2d(r, 2) + F.conv2d(F.avg_pool2d(xr, 2), wscr, bscr)
    else:
        r = F.conv2d(F.conv2d(F.relu(xr), w1r, b1r, padding=1) * m1, w2r, b2r, padding=1)
        s = F.conv2d(xr, wscr, bscr) if has_sc else xr
        if down:
            r, s = F.avg_pool2d(r, 2), F.avg_pool2d(s, 2)
        ref = r + s
    dy = torch.randn(ref.shape, generator=g)
    ref.backward(dy.double())
    xg = nhwc(x).to(dev).requires_grad_()
    ps = [t.to(dev).requires_grad_() for t in (w1, b1, w2, b2, wsc, bsc)]
    out = L.d_block_raw(xg, ps[0], ps[1], ps[2], ps[3], ps[4] if has_sc else None, ps[5] if has_sc else None,
                        down=down, optimized=optimized)
    out.backward(nhwc(dy).to(dev))
    sc = lambda t: max(t.abs().max().item(), 1.0)
    close(out.permute(0, 3, 1, 2), ref, 1e-3, 1e-4 * sc(ref), "fwd")
    close(xg.grad.permute(0, 3, 1, 2), xr.grad, 1e-3, 1e-4 * sc(xr.grad), "dx")
    names = ["w1", "b1", "w2", "b2", "wsc", "bsc"]
    for i, (p, rl) in enumerate(zip(ps, leaves[1:])):
        if i >= 4 and not has_sc:
            assert p.grad is None
            continue
        close(p.grad, rl.grad, 1e-3, 2e-4 * sc(rl.grad), "d" + names[i])


def test_psp_pool_and_concat_match_torch(dev):
    """csrc/psp.cu against the library composition of PSPModule (resnet_generator_app_v2.py:741-751): adaptive
    average pools (1,2,3,6), bilinear align_corners=True up-sampling, concat, 3x3 conv -- forward and backward."""
    from layout2img_b200 import functional as L
    g = torch.Generator().manual_seed(9)
    B, H, C, CP = 3, 64, 128, 100
    feats = torch.randn(B, C, H, H, generator=g)
    w = torch.randn(100, 4 * CP + C, 3, 3, generator=g) / (9 * (4 * CP + C)) ** 0.5
    mix = [torch.randn(CP, C, generator=g) / C ** 0.5 for _ in range(4)]
    fr, wr = feats.double().requires_grad_(), w.double().requires_grad_()
    priors = []
    for s, m in zip((1, 2, 3, 6), mix):
        p = F.relu(torch.einsum("oc,bchw->bohw", m.double(), F.adaptive_avg_pool2d(fr, s)))
        priors.append(F.interpolate(p, size=(H, H), mode="bilinear", align_corners=True))
    ref = F.conv2d(torch.cat(priors + [fr], 1), wr, None, padding=1)
    dy = torch.randn(ref.shape, generator=g)
    ref.backward(dy.double())

    fg, wg = nhwc(feats).to(dev).requires_grad_(), w.to(dev).requires_grad_()
    pooled = L.psp_pool(fg)
    ps, off = [], 0
    for s, m in zip((1, 2, 3, 6), mix):
        ps.append(F.relu(pooled[:, off:off + s * s] @ m.to(dev).t()))
        off += s * s
    out = L.psp_bottleneck(fg, torch.cat(ps, 1), wg)
    out.backward(nhwc(dy).to(dev))
    close(out.permute(0, 3, 1, 2), ref, 1e-3, 1e-4 * max(ref.abs().max().item(), 1), "psp fwd")
    close(fg.grad.permute(0, 3, 1, 2), fr.grad, 1e-3, 1e-4 * max(fr.grad.abs().max().item(), 1), "psp dfeats")
    close(wg.grad, wr.grad, 1e-3, 2e-4 * max(wr.grad.abs().max().item(), 1), "psp dw")


@pytest.mark.parametrize("shape,training", [((64, 3, 3, 3), True), ((1024, 512, 3, 3), True), ((256, 128, 1, 1), True),
                                            ((128, 64, 3, 3), False)])
def test_spectral_norm_kernels_match_torch(dev, shape, training):
    """csrc/specnorm.cu vs torch.nn.utils.spectral_norm: the in-place power iteration, sigma, the normalised
    weight (through the operand pair) and the gradient w.r.t. weight_orig."""
    from layout2img_b200 import ops
    g = torch.Generator().manual_seed(sum(shape))
    conv = torch.nn.Conv2d(shape[1], shape[0], shape[2], 1, shape[2] // 2)
    conv = torch.nn.utils.spectral_norm(conv, eps=1e-4)
    with torch.no_grad():
        conv.weight_orig.copy_(torch.randn(shape, generator=g) * 0.05)
    conv = conv.double()
    conv.train(training)
    w_orig = conv.weight_orig.detach().float().to(dev)
    u, v = conv.weight_u.detach().float().to(dev), conv.weight_v.detach().float().to(dev)
    for hook in conv._forward_pre_hooks.values():
        hook(conv, None)                                    # torch: power iteration (train) + W / sigma
    gy = torch.randn(shape, generator=g).double()
    (conv.weight * gy).sum().backward()
    st = ops.sn_sigma(w_orig, u, v, training, 1e-4)
    close(u, conv.weight_u, 1e-4, 1e-6, "u"); close(v, conv.weight_v, 1e-4, 1e-6, "v")
    close(st.u, conv.weight_u, 1e-4, 1e-6, "u used"); close(st.v, conv.weight_v, 1e-4, 1e-6, "v used")
    sigma_ref = torch.dot(conv.weight_u, conv.weight_orig.detach().reshape(shape[0], -1) @ conv.weight_v)
    close(st.sigma, sigma_ref.reshape(1), 1e-5, 1e-7, "sigma")
    taps = shape[2] * shape[3]
    wp = ops.conv_weight_prep(w_orig, st.sigma, need_dgrad=False)
    w_sn = (wp.f_hi.float() + wp.f_lo.float())[..., :shape[1]].reshape(shape[0], shape[2], shape[3], shape[1]).permute(0, 3, 1, 2)
    close(w_sn, conv.weight.detach(), 1e-4, 1e-5 * conv.weight.abs().max().item(), "W / sigma via the operand pair")
    g_ours = gy.float().permute(0, 2, 3, 1).reshape(shape[0], taps, shape[1]).contiguous().to(dev)
    dw = ops.sn_weight_grad(g_ours, w_orig, st)
    close(dw, conv.weight_orig.grad, 1e-4, 1e-5 * conv.weight_orig.grad.abs().max().item(), "d weight_orig")


@pytest.mark.parametrize("H,up2", [(4, True), (8, True), (16, False)])
def test_inorm_relu_upsample_matches_torch(dev, H, up2):
    """csrc/layout_ops.cu inorm_* vs InstanceNorm2d -> ReLU -> F.interpolate(bilinear, x2) of
    mask_regression.py:66-99 in fp64, forward (as the operand pair) and backward."""
    from layout2img_b200 import ops
    g = torch.Generator().manual_seed(H)
    N, C = 5, 256
    x = torch.randn(N, C, H, H, generator=g) * 2 + 0.5
    xr = x.double().requires_grad_()
    y = F.relu(F.instance_norm(xr, eps=1e-5))
    if up2:
        y = F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=False)
    da = torch.randn(y.shape, generator=g)
    y.backward(da.double())
    xg = nhwc(x).to(dev)
    pair, stats = ops.inorm_relu_fwd(xg, up2)
    got = (pair.hi.float() + pair.lo.float()).permute(0, 3, 1, 2)
    close(got, y, 1e-3, 1e-4, "inorm fwd pair")
    dx = ops.inorm_relu_bwd(xg, stats, nhwc(da).to(dev), up2)
    close(dx.permute(0, 3, 1, 2), xr.grad, 1e-3, 1e-4 * max(1.0, xr.grad.abs().max().item()), "inorm bwd")


def _sn_mod(mod, dev, seed):
    """Spectrally normalised module with deterministic weights and converged-ish u / v on `dev` (eval: no iteration)."""
    m = torch.nn.utils.spectral_norm(mod)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        m.weight_orig.copy_(torch.randn(m.weight_orig.shape, generator=g) * 0.2)
        if getattr(m, "bias", None) is not None:
            m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
    return m.to(dev)


@pytest.mark.parametrize("with_emb,training", [(False, True), (True, True), (True, False)])
def test_projection_head_matches_composition(dev, with_emb, training):
    """functional.proj_head (csrc/heads.cu) vs rcnn_discriminator_app.py:125-127 / :160-166 composed from torch ops in fp64
    (sum_hw relu -> spectral-norm linear [+ <spectral-norm embedding[y], .>]), forward, every gradient, u / v buffers."""
    import copy
    from layout2img_b200 import functional as L
    N, H, C, NC = 37, 4, 1024, 184
    g = torch.Generator().manual_seed(17)
    feat = torch.randn(N, H, H, C, generator=g)
    y = torch.randint(1, NC, (N,), generator=g)
    lin, emb = _sn_mod(torch.nn.Linear(C, 1), dev, 1), _sn_mod(torch.nn.Embedding(NC, C), dev, 2)
    lin_r, emb_r = copy.deepcopy(lin).cpu().double(), copy.deepcopy(emb).cpu().double()
    for m in (lin, emb, lin_r, emb_r):
        m.train(training)
    fr = feat.double().requires_grad_()
    s = F.relu(fr).sum(dim=(1, 2))
    ref = lin_r(s)
    if with_emb:
        ref = ref + (emb_r(y) * s).sum(dim=1, keepdim=True)
    dy = torch.randn(N, 1, generator=g)
    ref.backward(dy.double())
    fg = feat.to(dev).requires_grad_()
    out = L.proj_head(fg, lin, emb if with_emb else None, y.to(dev) if with_emb else None)
    out.backward(dy.to(dev))
    close(out, ref, what="head fwd")
    close(fg.grad, fr.grad, 1e-3, 1e-5, "head dfeat")
    close(lin.weight_orig.grad, lin_r.weight_orig.grad, 1e-3, 1e-4 * lin_r.weight_orig.grad.abs().max().item(), "head dw")
    close(lin.bias.grad, lin_r.bias.grad, 1e-3, 1e-4, "head dbias")
    close(lin.weight_u, lin_r.weight_u, 1e-4, 1e-5, "u")
    if with_emb:
        close(emb.weight_orig.grad, emb_r.weight_orig.grad, 1e-3, 1e-4 * emb_r.weight_orig.grad.abs().max().item(), "head demb")
        close(emb.weight_v, emb_r.weight_v, 1e-4, 1e-5, "emb v")


@pytest.mark.parametrize("training", [True, False])
def test_gram_projection_head_matches_reference_formula(dev, training):
    """functional.gram_proj (csrc/heads.cu) vs the reference's appearance head evaluated literally in fp64
    (rcnn_discriminator_app.py:148-157: Gram matrix, concat with the class embedding, linear, row mean)."""
    import copy
    from layout2img_b200 import functional as L
    K, H, C, NC = 9, 8, 64, 20
    g = torch.Generator().manual_seed(23)
    x = torch.randn(K, H, H, C, generator=g)
    y = torch.randint(1, NC, (K,), generator=g)
    app, emb = _sn_mod(torch.nn.Linear(2 * C, 1), dev, 3), _sn_mod(torch.nn.Embedding(NC, C), dev, 4)
    app_r, emb_r = copy.deepcopy(app).cpu().double(), copy.deepcopy(emb).cpu().double()
    for m in (app, emb, app_r, emb_r):
        m.train(training)
    xr = x.double().requires_grad_()
    f = F.relu(xr.permute(0, 3, 1, 2)).reshape(K, C, -1)               # (K, C, P) as the reference's NCHW view
    gram = torch.bmm(f, f.transpose(1, 2)) / C
    ey = emb_r(y)
    ref = app_r(torch.cat([gram, ey[:, None, :].expand(K, C, C)], dim=-1)).sum(1) / C
    dy = torch.randn(K, 1, generator=g)
    ref.backward(dy.double())
    xg = x.to(dev).requires_grad_()
    out = L.gram_proj(xg, app, emb, y.to(dev))
    out.backward(dy.to(dev))
    close(out, ref, what="gram fwd")
    close(xg.grad, xr.grad, 1e-3, 1e-4 * xr.grad.abs().max().item(), "gram dx")
    close(app.weight_orig.grad, app_r.weight_orig.grad, 1e-3, 1e-4 * app_r.weight_orig.grad.abs().max().item(), "gram dw")
    close(app.bias.grad, app_r.bias.grad, 1e-3, 1e-4, "gram dbias")
    close(emb.weight_orig.grad, emb_r.weight_orig.grad, 1e-3, 1e-4 * emb_r.weight_orig.grad.abs().max().item(), "gram demb")


@pytest.mark.parametrize("M,N,K,bias", [(512, 308, 308, True), (64, 16384, 128, True), (50, 100, 128, False), (7, 65, 33, True)])
def test_linear_kernels_match_torch(dev, M, N, K, bias):
    """functional.linear (a linear layer as a 1x1 tensor-core convolution over M one-pixel images: y = x W^T + b, dx = dy W,
    dW = dy^T x, db) vs fp64."""
    from layout2img_b200 import functional as L
    g = torch.Generator().manual_seed(M + N + K)
    x, w = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) if bias else None
    dy = torch.randn(M, N, generator=g)
    xr, wr = x.double().requires_grad_(), w.double().requires_grad_()
    br = b.double().requires_grad_() if bias else None
    ref = F.linear(xr, wr, br)
    ref.backward(dy.double())
    xg, wg = x.to(dev).requires_grad_(), w.to(dev).requires_grad_()
    bg = b.to(dev).requires_grad_() if bias else None
    out = L.linear(xg.view(1, M, K), wg, bg)                      # leading dimensions are flattened
    out.backward(dy.to(dev).view(1, M, N))
    close(out.view(M, N), ref, 1e-4, 2e-5 * ref.abs().max().item(), "linear fwd")
    close(xg.grad, xr.grad, 1e-4, 2e-5 * xr.grad.abs().max().item(), "linear dx")
    close(wg.grad, wr.grad, 1e-4, 2e-5 * wr.grad.abs().max().item(), "linear dw")
    if bias:
        close(bg.grad, br.grad, 1e-4, 2e-5 * br.grad.abs().max().item(), "linear db")


def test_add_layernorm_matches_torch(dev):
    """functional.add_layer_norm = LayerNorm(a + b) of the attention block (resnet_generator_app_v2.py:201-212)."""
    from layout2img_b200 import functional as L
    g = torch.Generator().manual_seed(8)
    rows, D = 37, 308
    a, b = torch.randn(3, rows, D, generator=g), torch.randn(3, rows, D, generator=g) * 0.5 + 0.2
    ln = torch.nn.LayerNorm(D)
    with torch.no_grad():
        ln.weight.copy_(torch.rand(D, generator=g) + 0.5); ln.bias.copy_(torch.randn(D, generator=g) * 0.1)
    import copy
    ln_r = copy.deepcopy(ln).double()
    dy = torch.randn(3, rows, D, generator=g)
    ar, br = a.double().requires_grad_(), b.double().requires_grad_()
    ref = ln_r(ar + br)
    ref.backward(dy.double())
    ln = ln.to(dev)
    ag, bg = a.to(dev).requires_grad_(), b.to(dev).requires_grad_()
    out = L.add_layer_norm(ag, bg, ln)
    out.backward(dy.to(dev))
    close(out, ref, 1e-4, 1e-5, "layernorm fwd")
    close(ag.grad, ar.grad, 1e-4, 1e-5, "layernorm da")
    close(bg.grad, br.grad, 1e-4, 1e-5, "layernorm db")
    close(ln.weight.grad, ln_r.weight.grad, 1e-4, 1e-4, "layernorm dweight")
    close(ln.bias.grad, ln_r.bias.grad, 1e-4, 1e-4, "layernorm dbias")


def test_maxpool2_matches_torch(dev):
    from layout2img_b200 import functional as L
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 16, 8, 8, generator=g).round(decimals=1)            # rounded: ties occur
    xr = x.double().requires_grad_()
    ref = F.max_pool2d(xr, 2)
    dy = torch.randn(ref.shape, generator=g)
    ref.backward(dy.double())
    xg = nhwc(x).to(dev).requires_grad_()
    out = L.maxpool2(xg)
    out.backward(nhwc(dy).to(dev))
    close(out.permute(0, 3, 1, 2), ref, 0, 0, "maxpool fwd")
    close(xg.grad.permute(0, 3, 1, 2), xr.grad, 0, 1e-12, "maxpool bwd (first maximum wins)")


@pytest.mark.parametrize("B,O,h,C", [(2, 4, 8, 100), (3, 8, 64, 100), (2, 16, 16, 100), (2, 31, 32, 36)])
def test_class_mix_matches_composition(dev, B, O, h, C):
    """functional.class_mix (gathered 1x1 mask head fused with the stage-mask mixing) vs the reference composition
    (resnet_generator_app_v2.py:646/651 + :466-470: full 184-channel 1x1 conv, gather, sigmoid, nearest / bilinear mixing)
    in fp64: forward and every gradient (features, conv weight / bias, alpha, bmask)."""
    from layout2img_b200 import functional as L
    NC, S = 184, 64
    g = torch.Generator().manual_seed(B * 1000 + O * 10 + h)
    t = torch.randn(B, C, h, h, generator=g)
    conv = torch.nn.Conv2d(C, NC, 1)
    alpha = torch.randn(1, NC, 1, generator=g) * 0.5
    bmask = torch.rand(B, O, S, S, generator=g)
    hard = (torch.rand(B, O, S, S, generator=g) > 0.4).float()
    y = torch.randint(0, NC, (B, O), generator=g)
    y[0, 1] = y[0, 0]                                           # two objects of one image share a class
    dy = torch.randn(B, O, h, h, generator=g)
    import copy
    conv_r = copy.deepcopy(conv).double()
    tr, ar, br = t.double().requires_grad_(), alpha.double().requires_grad_(), bmask.double().requires_grad_()
    stage = conv_r(tr)
    sel = torch.gather(stage, 1, y.view(B, O, 1, 1).expand(B, O, h, h))
    seman = torch.sigmoid(sel) * F.interpolate(hard.double(), size=(h, h), mode="nearest")
    a = torch.sigmoid(ar)[0, :, 0][y].view(B, O, 1, 1)
    soft = F.interpolate(br, size=(h, h), mode="bilinear", align_corners=False)
    ref = soft * (1 - a) + seman * a
    ref.backward(dy.double())
    conv = conv.to(dev)
    tg = nhwc(t).to(dev).requires_grad_()
    ag, bg = alpha.to(dev).requires_grad_(), bmask.to(dev).requires_grad_()
    out = L.class_mix(tg, conv, ag, bg, y.to(dev), hard.to(dev))
    out.backward(dy.to(dev))
    close(out, ref, 1e-4, 1e-5, "class_mix fwd")
    close(tg.grad.permute(0, 3, 1, 2), tr.grad, 1e-3, 1e-5, "d features")
    close(conv.weight.grad, conv_r.weight.grad, 1e-3, 1e-5 * max(1.0, conv_r.weight.grad.abs().max().item()), "d conv weight")
    close(conv.bias.grad, conv_r.bias.grad, 1e-3, 1e-5 * max(1.0, conv_r.bias.grad.abs().max().item()), "d conv bias")
    close(ag.grad, ar.grad, 1e-3, 1e-5 * max(1.0, ar.grad.abs().max().item()), "d alpha")
    close(bg.grad, br.grad, 1e-3, 1e-5, "d bmask")


@pytest.mark.gpu
def test_conv_tile_queue_variants_match():
    """The convolution kernel's tile order is a launch-time switch read once per process (L2I_CONV_DYNAMIC: tiles handed
    out by a device-wide counter instead of a static stride; L2I_CONV_NCAT=0: three-instruction product form).  Re-run
    the convolution parity tests in child processes with the non-default settings."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for var, val in (("L2I_CONV_DYNAMIC", "1"), ("L2I_CONV_NCAT", "0")):
        env = dict(os.environ, **{var: val})
        r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_ops.py"), "-q", "-x", "-k",
                            "test_conv_fwd_dgrad_wgrad or test_conv_epilogue_mask_pool_residual_pair"],
                           cwd=root, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, f"{var}={val}:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}"
