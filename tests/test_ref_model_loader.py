"""tools/ref_model.py: the UNMODIFIED reference model code (from /root/reference here, from the git-ignored bundle
oracle/_ref/aloception_src on the GPU box) imports with inert mocks for the third-party packages this image lacks, and
its operator is re-pointed at the B200 library by ``integration.install()``.  (The forward pass itself needs a GPU:
``tools/bench_model.py``; the reference asserts CUDA parameters, deformable_detr.py:254.)  Runs in a fresh interpreter:
other tests install namespace stubs for ``alonet`` that a full import must not meet."""
import os
import subprocess
import sys

import pytest

from tools import ref_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, %r)
import torch
from tools import ref_model
alonet, aloscene = ref_model.load()
import aloception_oss_b200 as msda
from alonet.deformable_detr import DeformableCriterion, DeformableDetrR50
import alonet.deformable_detr.ops.functions.ms_deform_attn_func as ref_func
import alonet.deformable_detr.ops.modules.ms_deform_attn as ref_mod

# the reference's loader (which would shell out to make.sh) is ours; its autograd Function and module are untouched
assert ref_func.load_MultiScaleDeformableAttention is msda.load_MultiScaleDeformableAttention
assert ref_mod.load_MultiScaleDeformableAttention is msda.load_MultiScaleDeformableAttention
assert ref_func.MSDeformAttnFunction is not msda.MSDeformAttnFunction
model = DeformableDetrR50(num_classes=91, device=torch.device("cpu"))
attn = [m for m in model.modules() if type(m).__name__ == "MSDeformAttn"]
assert len(attn) == 12 and all(type(m).__module__ == "alonet.deformable_detr.ops.modules.ms_deform_attn" for m in attn)
n = sum(p.numel() for p in model.parameters())
assert 39e6 < n < 41e6, n
assert hasattr(torch.ops.alonet_custom, "ms_deform_attn_forward")
q = torch.randn(1, 5, 256)
shapes = torch.tensor([[4, 4], [2, 2], [1, 1], [1, 1]], dtype=torch.int32)
start = torch.tensor([0, 16, 20, 21], dtype=torch.int32)
try:  # CPU tensors: the reference's own error, raised by OUR registration of the op
    attn[0](q, torch.rand(1, 5, 4, 2), torch.randn(1, 22, 256), shapes, start)
    raise SystemExit("CPU call did not raise")
except RuntimeError as e:
    assert "Not implemented on the CPU" in str(e), str(e)
f = aloscene.Frame(torch.rand(3, 32, 48), names=("C", "H", "W")).norm_resnet()
batch = aloscene.Frame.batch_list([f, f])
assert tuple(batch.shape) == (2, 3, 32, 48) and batch.mask is not None
print("REF MODEL LOADER OK", len(attn), round(n / 1e6, 2))
''' % ROOT


@pytest.mark.skipif(ref_model.source_root() is None, reason="reference model sources not available")
def test_reference_detr_r50_builds_on_the_b200_operator():
    res = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert "REF MODEL LOADER OK 12 40.07" in res.stdout
