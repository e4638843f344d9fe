"""Recipe: bundle the reference's PYTHON model sources for the model-level benchmark on the GPU box.

BASELINE.json configs[3] / configs[4] are quoted on the reference's own DeformableDETR-R50 (alonet/deformable_detr/
deformable_detr_r50.py:19-30, deformable_detr.py:215-299, criterion / matcher, DDP harness).  That model is NOT rebuilt here
(everything above the MSDeformAttn module stays): ``tools/bench_model.py`` runs the UNMODIFIED reference
classes with ``aloception_oss_b200.integration.install()`` pointing their operator at the B200 kernels.  /root/reference does
not exist on the GPU box, so -- like oracle/_ref/*.so -- this script copies the ``*.py`` files of ``alonet`` and ``aloscene``
byte for byte into the git-ignored ``oracle/_ref/aloception_src/`` (in .gitignore, not in .gpurunignore); nothing of it enters the
repository's history.  Run by ``__graft_entry__.build()`` when /root/reference is present.

MEASUREMENT INFRASTRUCTURE ONLY; never imported by the product package.
"""
from __future__ import annotations

import os
import shutil

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("MSDA_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(_HERE, "_ref", "aloception_src")
PACKAGES = ("alonet", "aloscene", "alodataset")


def reference_available() -> bool:
    return all(os.path.isdir(os.path.join(REFERENCE_ROOT, p)) for p in PACKAGES)


def bundled() -> bool:
    return os.path.isfile(os.path.join(DST, "alonet", "deformable_detr", "deformable_detr_r50.py"))


def build() -> str:
    n = 0
    for pkg in PACKAGES:
        src_root = os.path.join(REFERENCE_ROOT, pkg)
        for dirpath, dirnames, filenames in os.walk(src_root):
            dirnames[:] = [d for d in dirnames if d not in ("__pycache__", "build")]
            rel = os.path.relpath(dirpath, REFERENCE_ROOT)
            for fn in filenames:
                if fn.endswith(".py"):
                    os.makedirs(os.path.join(DST, rel), exist_ok=True)
                    shutil.copyfile(os.path.join(dirpath, fn), os.path.join(DST, rel, fn))
                    n += 1
    return f"{DST} ({n} python files)"


if __name__ == "__main__":
    print(build() if reference_available() else "reference tree not present")
