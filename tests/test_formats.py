"""The on-disk / loader formats either side of the hot path, on CPU: checkpoint files in the
reference authors' format (train_context_app_v2.py:77-103,215-217) and the dataset item contract
(data/cocostuff_loader.py:222-380,301-303).  Also the drop-in import path (train_context_app_v2.py:18-19)."""
import os

import numpy as np
import pytest
import torch

from conftest import load_schema
from layout2img_b200.synth import make_state


def _G(num_classes=184):
    from layout2img_b200.model.resnet_generator_app_v2 import ResnetGenerator128_context
    return ResnetGenerator128_context(num_classes=num_classes, output_dim=3)


def test_checkpoint_round_trip_module_prefix(tmp_path):
    from layout2img_b200.checkpoint import load_checkpoint, save_checkpoint
    G = _G()
    G.load_state_dict(make_state(load_schema("G"), 5))
    path = os.path.join(tmp_path, "G_5.pth")
    save_checkpoint(G, path)
    raw = torch.load(path)
    assert all(k.startswith("module.") for k in raw) and len(raw) == len(G.state_dict()) == 281
    # a DataParallel-wrapped torch module writes exactly these keys
    assert list(raw) == list(torch.nn.DataParallel(G).state_dict())
    G2 = _G()
    rep = load_checkpoint(G2, path)
    assert not rep["ignored"] and not rep["missing"] and len(rep["loaded"]) == 281
    for k, v in G.state_dict().items():
        assert torch.equal(v, G2.state_dict()[k]), k


def test_checkpoint_intersects_keys_like_the_reference(tmp_path):
    """train_context_app_v2.py:86-88: keys the model does not know are dropped, keys the file lacks keep their init."""
    from layout2img_b200.checkpoint import load_checkpoint
    from layout2img_b200.model.rcnn_discriminator_app import CombineDiscriminator128_app
    D = CombineDiscriminator128_app(num_classes=184)
    sd = {"module." + k: torch.full_like(v, 0.5) if v.is_floating_point() else v for k, v in D.state_dict().items()}
    dropped = "module.obD.l7.bias"
    del sd[dropped]
    sd["module.obD.some_removed_head.weight"] = torch.zeros(3)
    path = os.path.join(tmp_path, "D_5.pth")
    torch.save(sd, path)
    D2 = CombineDiscriminator128_app(num_classes=184)
    before = D2.state_dict()["obD.l7.bias"].clone()
    rep = load_checkpoint(D2, path)
    assert rep["ignored"] == ["obD.some_removed_head.weight"] and rep["missing"] == ["obD.l7.bias"]
    assert torch.equal(D2.state_dict()["obD.l7.bias"], before)
    assert float(D2.state_dict()["obD.block1.conv1.weight_orig"].mean()) == 0.5
    # a file written without the DataParallel prefix loads too
    torch.save({k[7:]: v for k, v in sd.items()}, path)
    rep = load_checkpoint(CombineDiscriminator128_app(num_classes=184), path)
    assert len(rep["loaded"]) == 129


def test_pack_layout_pads_like_the_loader():
    from layout2img_b200.data import PAD_BOX, collate, pack_layout
    img = np.zeros((3, 128, 128), dtype=np.float32)
    image, objs, boxes = pack_layout(img, [7, 183, 12], [[0.1, 0.2, 0.3, 0.4], [0, 0, 1, 1], [0.5, 0.5, 0.25, 0.25]], 8)
    assert image.shape == (3, 128, 128) and objs.dtype == torch.int64 and boxes.dtype == torch.float32
    assert objs.tolist() == [7, 183, 12, 0, 0, 0, 0, 0]
    assert boxes.shape == (8, 4) and torch.allclose(boxes[3:], torch.tensor(PAD_BOX).expand(5, 4))
    assert torch.allclose(boxes[0], torch.tensor([0.1, 0.2, 0.3, 0.4]))
    with pytest.raises(ValueError):
        pack_layout(img, list(range(9)), [[0, 0, 1, 1]] * 9, 8)
    real, label, bbox = collate([pack_layout(img, [1], [[0, 0, 1, 1]], 8), pack_layout(img, [2, 3], [[0, 0, 1, 1]] * 2, 8)])
    assert real.shape == (2, 3, 128, 128) and label.shape == (2, 8) and bbox.shape == (2, 8, 4)
    # the same padding entry the synthetic layouts use
    from layout2img_b200 import synth
    assert tuple(synth.PAD_BOX) == tuple(PAD_BOX)


def test_drop_in_import_path():
    """`from model.resnet_generator_app_v2 import *` / `from model.rcnn_discriminator_app import *` (the reference's
    train_context_app_v2.py:18-19, test_context_app_v2.py:10) resolve to the B200-native modules unchanged."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("from model.resnet_generator_app_v2 import *\nfrom model.rcnn_discriminator_app import *\n"
            "from model.sync_batchnorm import DataParallelWithCallback\n"
            "G = ResnetGenerator128_context(num_classes=184, output_dim=3)\nD = CombineDiscriminator128_app(num_classes=184)\n"
            "assert type(G).__module__.startswith('layout2img_b200.') and len(G.state_dict()) == 281 and len(D.state_dict()) == 130\n"
            "assert DataParallelWithCallback(G) is G\n"
            "for n in ('conv2d', 'bbox_mask', 'batched_index_select', 'BatchNorm'): assert n in globals(), n\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]


def test_prepared_weight_entries_are_tied_to_the_live_unmodified_tensor():
    """sn_group.PREPARED hands the grouped spectral-norm results to the autograd nodes keyed by id(weight); an entry must
    not be honoured for another tensor that recycled the id, nor after the weight was modified (optimizer step)."""
    import weakref
    from layout2img_b200 import sn_group
    w = torch.nn.Parameter(torch.ones(4, 4))
    sn_group.PREPARED[id(w)] = (weakref.ref(w), w._version, "state", "pairs")
    assert sn_group.has_prepared(w)
    assert sn_group.take_prepared(w) == ("state", "pairs") and not sn_group.has_prepared(w)     # consumed once
    sn_group.PREPARED[id(w)] = (weakref.ref(w), w._version, "state", "pairs")
    with torch.no_grad():
        w.add_(1.0)                                                                              # version counter moves
    assert not sn_group.has_prepared(w) and sn_group.take_prepared(w) is None
    other = torch.nn.Parameter(torch.zeros(4, 4))
    sn_group.PREPARED[id(other)] = (weakref.ref(w), other._version, "state of w", "pairs of w")  # a recycled id
    assert not sn_group.has_prepared(other) and sn_group.take_prepared(other) is None
    assert id(other) not in sn_group.PREPARED
