"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/l2i.h declares; module surfaces carry the reference's state_dict schema; no compute."""
import ctypes
import os

import pytest
import torch

from conftest import ROOT, load_schema
from layout2img_b200 import _lib


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = _lib.declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/l2i.h but not exported by libl2i.so"
    assert lib.l2i_version() >= 100
    assert set(_lib.prototypes()) == set(names)


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "layout2img_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), f"{f} must not reference the test oracle"


@pytest.mark.parametrize("kind,cls_path", [("G", "resnet_generator_app_v2.ResnetGenerator128_context"),
                                           ("G_nocontext", "resnet_generator_app_v2.ResnetGenerator128"),
                                           ("D", "rcnn_discriminator_app.CombineDiscriminator128_app")])
def test_state_dict_schema_matches_reference(kind, cls_path):
    import importlib
    mod, cls = cls_path.split(".")
    m = importlib.import_module("layout2img_b200.model." + mod)
    net = getattr(m, cls)(num_classes=184)
    sch = load_schema(kind)
    sd = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert list(sd) == list(sch)
    assert sd == dict(sch)


def test_modules_refuse_cpu_tensors():
    from layout2img_b200.model.rcnn_discriminator_app import CombineDiscriminator128_app
    D = CombineDiscriminator128_app(num_classes=10)
    with pytest.raises(RuntimeError, match="CUDA"):
        D(torch.zeros(1, 3, 128, 128), torch.zeros(1, 2, 4), torch.ones(1, 2, dtype=torch.long))
