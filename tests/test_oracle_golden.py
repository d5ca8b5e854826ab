"""Pins oracle/l2i_oracle.py against outputs of the unmodified reference (tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

from conftest import assert_summary_close, load_case, load_schema, dropout_keep_mask
from layout2img_b200.synth import make_state, synthetic_layout
from oracle import l2i_oracle as O

RTOL, ATOL = 1e-3, 1e-4      # north_star tolerance; the oracle is in practice ~1e-6 from the reference


def _setup(name):
    z, meta = load_case(name)
    data = synthetic_layout(meta["batch"], meta["num_obj"], meta["num_classes"], seed=meta["seed"],
                            n_pad=meta["n_pad"])
    PG = make_state(load_schema("G", meta["num_classes"]), meta["seed_g"])
    PD = make_state(load_schema("D", meta["num_classes"]), meta["seed_d"])
    return z, meta, data, PG, PD


@pytest.mark.parametrize("name", ["C", "Cpad", "V"])
def test_leaf_ops(name):
    z, meta, data, PG, PD = _setup(name)
    bbox = data["bbox"]
    for size in (64, 128):
        got = np.packbits(O.bbox_mask(bbox, size, size).numpy().astype(np.uint8))
        assert np.array_equal(got, z[f"op.bbox_mask{size}"]), "bbox_mask must be bit-exact"
    np.testing.assert_allclose(O.box_relational_embedding(bbox).numpy(), z["op.box_rel_emb"], rtol=1e-5, atol=1e-5)
    gm = torch.Generator().manual_seed(meta["seed"] + 100)
    masks = torch.rand(meta["batch"], meta["num_obj"], 16, 16, generator=gm)
    np.testing.assert_allclose(O.masks_to_layout(bbox, masks, 64).numpy(), z["op.masks_to_layout"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", ["C", "Cpad", "V"])
def test_eval_forward(name):
    z, meta, data, PG, PD = _setup(name)
    taps = {}
    with torch.no_grad():
        fake = O.g_forward(PG, data["z"], data["bbox"], data["z_im"], data["label"], False, taps=taps)
        d_out = O.d_forward(PD, fake, data["bbox"], data["label"], False)
    if name == "C":
        np.testing.assert_allclose(fake.numpy(), z["eval.fake"], rtol=RTOL, atol=ATOL)
    assert_summary_close(fake, z["eval.fake.sum"], RTOL, ATOL, "fake")
    assert_summary_close(taps["bmask"], z["eval.bmask"], RTOL, ATOL, "bmask")
    for k in range(1, 6):
        assert_summary_close(taps[f"x{k}"], z[f"eval.x{k}"], RTOL, ATOL, f"x{k}")
    for i, nm in enumerate(("d_im", "d_obj", "d_app")):
        np.testing.assert_allclose(d_out[i].numpy(), z[f"eval.{nm}"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("name", ["C", "Cpad", "V"])
def test_train_step(name):
    """One full D+G iteration: losses, every gradient, every post-step parameter and buffer."""
    z, meta, data, PG, PD = _setup(name)
    O.set_requires_grad(PG); O.set_requires_grad(PD)
    g_opt, d_opt = O.make_adam(PG, 1e-4), O.make_adam(PD, 1e-4)
    grads = {}
    # record D grads right after d_loss.backward (the G step pollutes them afterwards)
    orig_step = d_opt.step
    def step_and_record(*a, **k):
        for n in O.param_names(PD):
            grads["d." + n] = PD[n].grad.detach().clone()
        return orig_step(*a, **k)
    d_opt.step = step_and_record
    torch.manual_seed(meta["dropout_seed"])
    d_loss, g_loss, fake = O.train_step(PG, PD, g_opt, d_opt, data["real"], data["label"], data["bbox"],
                                        data["z"], data["z_im"])
    assert abs(d_loss.item() - float(z["train.d_loss"])) < 1e-4
    assert abs(g_loss.item() - float(z["train.g_loss"])) < 1e-4
    assert_summary_close(fake, z["train.fake"], RTOL, ATOL, "train fake")
    for n in O.param_names(PD):
        assert_summary_close(grads["d." + n], z[f"train.dgrad.{n}"], RTOL, ATOL, "dgrad " + n)
    for n in O.param_names(PG):
        want = z[f"train.ggrad.{n}"]
        # d(mask) carries the reference's 1/(sum_o m + 1e-6) amplification
        assert_summary_close(PG[n].grad, want, 2e-3, 1e-4 + 1e-5 * abs(want[2]), "ggrad " + n)
    for n, v in PG.items():
        assert_summary_close(v.float(), z[f"train.gstate.{n}"], RTOL, ATOL, "gstate " + n)
    for n, v in PD.items():
        assert_summary_close(v.float(), z[f"train.dstate.{n}"], RTOL, ATOL, "dstate " + n)


def test_dropout_mask_matches_reference_draw():
    """An explicit keep-mask reproduces the seeded F.dropout2d draw the goldens used."""
    z, meta, data, PG, PD = _setup("C")
    mask = dropout_keep_mask(meta["dropout_seed"], meta["batch"])
    with torch.no_grad():
        a = O.g_forward({k: v.clone() for k, v in PG.items()}, data["z"], data["bbox"], data["z_im"],
                        data["label"], True, dropout_mask=mask)
        torch.manual_seed(meta["dropout_seed"])
        b = O.g_forward({k: v.clone() for k, v in PG.items()}, data["z"], data["bbox"], data["z_im"],
                        data["label"], True)
    assert torch.equal(a, b)
