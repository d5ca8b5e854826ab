import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# libl2i.so is a build artefact (git-ignored): make sure it exists and is current before any test imports it.
# nvcc cross-compiles sm_100a without a GPU; a no-op when the library is newer than its sources.
try:
    from layout2img_b200 import build as _l2i_build
    _l2i_build.build()
except Exception as _e:          # the ABI tests will then fail with the loader's own message
    sys.stderr.write(f"conftest: building libl2i.so failed: {_e}\n")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_schema(kind: str, num_classes: int = 184):
    """Reference state_dict schema (name -> shape) recorded by tests/golden/make_golden.py."""
    sch = json.load(open(os.path.join(GOLDEN, f"schema_{kind}.json")))
    out = {}
    for k, v in sch.items():
        v = list(v)
        if num_classes != 184 and (k == "label_embedding.weight" or ".l_y." in k or ".l_y_app." in k):
            if v and v[0] == 184 and not k.endswith("weight_v"):
                v[0] = num_classes
        out[k] = tuple(v)
    return out


def load_case(name: str):
    z = np.load(os.path.join(GOLDEN, f"case_{name}.npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


def summarize(t: torch.Tensor, n: int = 64) -> np.ndarray:
    """Same summary as tests/golden/make_golden.py: [sum, abs-sum, l2, n strided samples]."""
    f = t.detach().double().cpu().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, steps=min(n, f.numel())).long()
    head = torch.stack([f.sum(), f.abs().sum(), f.norm()])
    samp = torch.zeros(n, dtype=torch.float64)
    samp[: idx.numel()] = f[idx]
    return torch.cat([head, samp]).numpy()


def assert_summary_close(got: torch.Tensor, want: np.ndarray, rtol: float, atol: float, what: str = ""):
    """Compare a tensor with a stored summary: samples element-wise at (rtol, atol); the
    three global moments relative to the l2 norm (they aggregate numel rounding errors)."""
    n = want.shape[0] - 3
    g = summarize(got, n)
    np.testing.assert_allclose(g[3:], want[3:], rtol=rtol, atol=atol, err_msg=f"{what}: samples")
    l2 = max(abs(want[2]), 1e-12)
    scale = np.sqrt(max(got.numel(), 1))
    assert abs(g[2] - want[2]) <= rtol * l2 + atol * scale, f"{what}: l2 {g[2]} vs {want[2]}"
    assert abs(g[1] - want[1]) <= rtol * abs(want[1]) + atol * got.numel(), f"{what}: abs-sum"
    assert abs(g[0] - want[0]) <= rtol * abs(want[1]) + atol * got.numel(), f"{what}: sum"


def dropout_keep_mask(seed: int, batch: int, size: int = 64, p: float = 0.1) -> torch.Tensor:
    """The (b,100,1,1) scaled keep-mask the reference's PSP Dropout2d draws when the CPU RNG
    is seeded with ``seed`` right before the step (make_golden.py does exactly that)."""
    torch.manual_seed(seed)
    m = torch.nn.functional.dropout2d(torch.ones(batch, 100, size, size), p, True)
    return m[:, :, :1, :1].clone()
