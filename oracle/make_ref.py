"""Stage the UNMODIFIED reference modules of the hot path under baseline/_ref/.  TEST / BENCH INFRASTRUCTURE ONLY.

    python oracle/make_ref.py            (build container only: needs /root/reference)

The reference is pure Python (no native sources, setup.py globs an absent model/csrc), so
"building" it is copying the files the path imports byte for byte into the git-ignored,
gpurun-shipped baseline/_ref/.  Nothing is copied into the repository's history and nothing under
layout2img_b200/ may import from there.  Users:
  * bench.py --impl reference       the reference's own CPU path, timed on the box's host cores
  * bench.py (library_baseline)     the same modules under PyTorch eager + cuDNN/cuBLAS on the B200
  * tests/test_reference_gpu.py     the CUDA path against the reference modules on the same GPU
`pip install /root/reference` is not applicable (the reference's setup.py only builds the absent extension).
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")

FILES = [
    "model/__init__.py", "model/resnet_generator_app_v2.py", "model/rcnn_discriminator_app.py",
    "model/norm_module.py", "model/mask_regression.py",
    "model/sync_batchnorm/__init__.py", "model/sync_batchnorm/batchnorm.py", "model/sync_batchnorm/comm.py",
    "model/sync_batchnorm/replicate.py", "model/sync_batchnorm/batchnorm_reimpl.py", "model/sync_batchnorm/unittest.py",
    "utils/__init__.py", "utils/bilinear.py", "utils/util.py",
]


def stage(verbose: bool = False) -> bool:
    """Copy the files when /root/reference exists; returns True when baseline/_ref is complete afterwards."""
    if os.path.isdir(SRC):
        for rel in FILES:
            s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            if not os.path.exists(d) or not filecmp.cmp(s, d, shallow=False):
                shutil.copyfile(s, d)
                if verbose:
                    print("staged", rel)
    return all(os.path.exists(os.path.join(DST, rel)) for rel in FILES)


def available() -> bool:
    return all(os.path.exists(os.path.join(DST, rel)) for rel in FILES)


if __name__ == "__main__":
    ok = stage(verbose=True)
    print("baseline/_ref", "complete" if ok else "INCOMPLETE")
    sys.exit(0 if ok else 1)
