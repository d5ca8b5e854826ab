"""End-to-end parity on the B200: the drop-in G and D modules (CUDA path through libl2i.so) against
the CPU oracle on the same seeded inputs and weights, and against the reference-generated golden
summaries.  Forward (eval), one full training iteration (losses, every gradient, post-step buffers)."""
import numpy as np
import pytest
import torch

from conftest import assert_summary_close, dropout_keep_mask, load_case, load_schema
from layout2img_b200.synth import make_state, synthetic_layout
from oracle import l2i_oracle as O

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-4          # north_star tolerance (fp32)


def _build(meta, dev):
    from layout2img_b200.model.rcnn_discriminator_app import CombineDiscriminator128_app
    from layout2img_b200.model.resnet_generator_app_v2 import ResnetGenerator128_context
    PG = make_state(load_schema("G", meta["num_classes"]), meta["seed_g"])
    PD = make_state(load_schema("D", meta["num_classes"]), meta["seed_d"])
    G = ResnetGenerator128_context(num_classes=meta["num_classes"], output_dim=3)
    D = CombineDiscriminator128_app(num_classes=meta["num_classes"])
    G.load_state_dict(PG)
    D.load_state_dict(PD)
    return G.to(dev), D.to(dev), PG, PD


def _data(meta):
    return synthetic_layout(meta["batch"], meta["num_obj"], meta["num_classes"], seed=meta["seed"], n_pad=meta["n_pad"])


def close(got, want, rtol=RTOL, atol=ATOL, what=""):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    err = (got - want).abs()
    bad = err > atol + rtol * want.abs()
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} outside tol, max err {err.max().item():.3e} (ref max {want.abs().max().item():.3e})"


@pytest.mark.parametrize("name", ["C", "Cpad", "V"])
def test_eval_forward_matches_oracle_and_reference(name):
    dev = torch.device("cuda:0")
    z, meta = load_case(name)
    data = _data(meta)
    G, D, PG, PD = _build(meta, dev)
    G.eval(); D.eval()
    with torch.no_grad():
        fake = G(data["z"].to(dev), data["bbox"].to(dev), data["z_im"].to(dev), data["label"].to(dev))
        d_out = D(fake, data["bbox"].to(dev), data["label"].to(dev).unsqueeze(-1))
        ref_fake = O.g_forward(PG, data["z"], data["bbox"], data["z_im"], data["label"], False)
        ref_d = O.d_forward(PD, ref_fake, data["bbox"], data["label"], False)
    assert fake.shape == (meta["batch"], 3, 128, 128)
    close(fake, ref_fake, what="G eval forward vs oracle")
    assert_summary_close(fake.cpu(), z["eval.fake.sum"], RTOL, ATOL, "G eval forward vs reference golden")
    for i, nm in enumerate(("d_im", "d_obj", "d_app")):
        close(d_out[i], ref_d[i], what=nm + " vs oracle")
        np.testing.assert_allclose(d_out[i].cpu().numpy(), z[f"eval.{nm}"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("name", ["C", "Cpad", "V"])
def test_train_step_matches_oracle_and_reference(name):
    """One D+G iteration: losses, D grads after d_loss.backward, G grads after g_loss.backward, and the
    spectral-norm / batch-norm buffers + Adam-updated parameters afterwards."""
    from layout2img_b200.train import make_optimizers, train_step
    dev = torch.device("cuda:0")
    z, meta = load_case(name)
    data = _data(meta)
    G, D, PG, PD = _build(meta, dev)
    G.train(); D.train()
    keep = dropout_keep_mask(meta["dropout_seed"], meta["batch"])          # (b,100,1,1)
    G.res4.conv_mask[0].dropout_mask = keep.view(meta["batch"], 100)
    g_opt, d_opt = make_optimizers(G, D)
    grads = {}

    def record(tag):
        net = D if tag == "d" else G
        for n, p in net.named_parameters():
            grads[tag + "." + n] = p.grad.detach().clone()

    d_loss, g_loss, fake = train_step(G, D, g_opt, d_opt, data["real"].to(dev), data["label"].to(dev),
                                      data["bbox"].to(dev), data["z"].to(dev), data["z_im"].to(dev), record=record)
    # ---- oracle on CPU
    O.set_requires_grad(PG); O.set_requires_grad(PD)
    og, od = O.make_adam(PG, 1e-4), O.make_adam(PD, 1e-4)
    ref = {}
    od_step = od.step
    def d_step(*a, **k):
        for n in O.param_names(PD):
            ref["d." + n] = PD[n].grad.detach().clone()
        return od_step(*a, **k)
    od.step = d_step
    og_step = og.step
    def g_step(*a, **k):
        for n in O.param_names(PG):
            ref["g." + n] = PG[n].grad.detach().clone()
        return og_step(*a, **k)
    og.step = g_step
    rd, rg, rfake = O.train_step(PG, PD, og, od, data["real"], data["label"], data["bbox"], data["z"], data["z_im"],
                                 dropout_mask=keep)
    assert abs(d_loss.item() - rd.item()) < 1e-4 and abs(d_loss.item() - float(z["train.d_loss"])) < 1e-4
    assert abs(g_loss.item() - rg.item()) < 1e-4 and abs(g_loss.item() - float(z["train.g_loss"])) < 1e-4
    close(fake, rfake, what="train fake")
    assert_summary_close(fake.cpu(), z["train.fake"], RTOL, ATOL, "train fake vs reference golden")
    for n in O.param_names(PD):
        want = ref["d." + n]
        close(grads["d." + n], want, 2e-3, 1e-5 + 2e-4 * want.abs().max().item(), "D grad " + n)
        assert_summary_close(grads["d." + n].cpu(), z[f"train.dgrad.{n}"], 2e-3, 1e-5 + 2e-4 * want.abs().max().item(), "D grad vs golden " + n)
    for n in O.param_names(PG):
        want = ref["g." + n]
        # d(mask) carries the reference's 1/(sum_o m + 1e-6) amplification
        close(grads["g." + n], want, 2e-3, 1e-5 + 5e-4 * want.abs().max().item(), "G grad " + n)
    sdG, sdD = G.state_dict(), D.state_dict()
    for n, v in PG.items():
        close(sdG[n].float(), v.float(), 1e-3, 2e-4, "G state " + n)
    for n, v in PD.items():
        close(sdD[n].float(), v.float(), 1e-3, 2e-4, "D state " + n)


def test_train_mode_buffers_advance_like_reference():
    """block_obj4 is shared by both object branches: its spectral-norm u/v must iterate twice per forward."""
    dev = torch.device("cuda:0")
    z, meta = load_case("C")
    data = _data(meta)
    G, D, PG, PD = _build(meta, dev)
    D.train()
    with torch.no_grad():
        D(data["real"].to(dev), data["bbox"].to(dev), data["label"].to(dev).unsqueeze(-1))
        O.d_forward(PD, data["real"], data["bbox"], data["label"], True)
    sd = D.state_dict()
    for n in ("obD.block_obj4.conv1.weight_u", "obD.block_obj4.conv2.weight_v", "obD.block1.conv1.weight_u"):
        close(sd[n], PD[n], 1e-4, 1e-5, n)


def test_bbox_on_cpu_and_positional_call():
    """The reference's loop keeps bbox on the CPU (train_context_app_v2.py:153) and the samplers call
    netG.forward(z_obj, bbox, z_im, label) positionally (test_context_app_v2.py:77)."""
    dev = torch.device("cuda:0")
    z, meta = load_case("C")
    data = _data(meta)
    G, D, _, _ = _build(meta, dev)
    G.eval(); D.eval()
    with torch.no_grad():
        bb = data["bbox"].clone()
        a = G.forward(data["z"].to(dev), bb, data["z_im"].to(dev), data["label"].to(dev))
        b = G(data["z"].to(dev), bb.to(dev), data["z_im"].to(dev), y=data["label"].to(dev))
        D(a, bb, data["label"].to(dev))
    assert torch.equal(a, b)
    assert torch.equal(bb, data["bbox"]), "bbox must not be mutated"
