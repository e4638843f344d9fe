"""Golden vectors of the reference MODULE ``MSDeformAttn.forward`` (run in the build container only).

    python oracle/make_golden_module.py

TEST INFRASTRUCTURE.  Instantiates the UNMODIFIED reference class
(alonet/deformable_detr/ops/modules/ms_deform_attn.py:34-155, loaded through
``aloception_oss_b200.integration.import_reference_ops``), loads the deterministic weights of
``aloception_oss_b200.synthetic.module_case`` and evaluates forward (pure-PyTorch branch, ``is_tracing``) and autograd
backward in float64.  Stored: output, gradients w.r.t. query / input_flatten / reference_points and w.r.t. every
parameter.  These pin the FUSED operator (softmax + location arithmetic inside the kernels).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from aloception_oss_b200 import integration  # noqa: E402
from aloception_oss_b200.synthetic import MODULE_CASES, module_case  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden_module")


def main():
    os.makedirs(GOLD, exist_ok=True)
    _, modules = integration.import_reference_ops(os.path.join(ref_loader.REFERENCE_ROOT, "alonet"))
    for name in MODULE_CASES:
        cfg, state, x = module_case(name)
        mod = modules.MSDeformAttn(cfg["d_model"], cfg["n_levels"], cfg["n_heads"], cfg["n_points"]).double()
        mod.load_state_dict({k: torch.from_numpy(v).double() for k, v in state.items()})
        q = torch.from_numpy(x["query"]).double().requires_grad_(True)
        ref = torch.from_numpy(x["reference_points"]).double().requires_grad_(True)
        src = torch.from_numpy(x["input_flatten"]).double().requires_grad_(True)
        mask = None if x["mask"] is None else torch.from_numpy(x["mask"])
        # the reference's zero-padding helper allocates float32 zeros; run the module in float64 needs float64 pads, so
        # evaluate in float64 only the parts autograd needs: cast happens inside torch.cat (type promotion) -> fine
        out = mod(q, ref, src, torch.from_numpy(x["shapes"]), torch.from_numpy(x["start"]), mask, is_tracing=None)
        out.backward(torch.from_numpy(x["grad_out"]).double())
        rec = {"name": name, "out": out.detach().numpy(), "g_query": q.grad.numpy(), "g_ref": ref.grad.numpy(),
               "g_src": src.grad.numpy()}
        for k, p in mod.named_parameters():
            rec["gp_" + k] = p.grad.numpy()
        path = os.path.join(GOLD, name + ".npz")
        np.savez_compressed(path, **{k: (v.astype(np.float32) if isinstance(v, np.ndarray) else v) for k, v in rec.items()})
        print(f"{name}: out{tuple(out.shape)} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
