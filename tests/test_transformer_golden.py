"""Encoder / decoder layer loop (aloception_oss_b200/transformer.py) against golden vectors of the
UNMODIFIED reference classes (tests/golden_transformer/, produced by oracle/make_golden_transformer.py from
alonet/deformable_detr/deformable_transformer.py:306-632).

CPU: our mirror on the tracing (pure-PyTorch operator) branch in float64 loads the reference's state_dict and reproduces
outputs and input gradients -- pins the mirror and the fixtures.
GPU: the same modules through the CUDA operator (fused and unfused), float32, eager and CUDA-graph replay.
"""
import os

import numpy as np
import pytest
import torch

from aloception_oss_b200 import transformer as T
from tests._util import assert_close, rms

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_transformer")
LEVELS = ((12, 16), (6, 8), (3, 4), (2, 2))
D_MODEL, D_FFN, HEADS, POINTS = 64, 128, 2, 4
CASES = ("encoder2", "decoder2_ref2", "decoder2_ref4_refine")


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd_")}
    return z, sd


def build(name, sd, device, dtype, fused):
    if name.startswith("encoder"):
        m = T.DeformableTransformerEncoder(
            T.DeformableTransformerEncoderLayer(D_MODEL, D_FFN, 0.0, "relu", len(LEVELS), HEADS, POINTS, fused=fused), 2)
    else:
        m = T.DeformableTransformerDecoder(
            T.DeformableTransformerDecoderLayer(D_MODEL, D_FFN, 0.0, "relu", len(LEVELS), HEADS, POINTS, fused=fused), 2,
            return_intermediate=True)
        if "refine" in name:
            m.bbox_embed = torch.nn.ModuleList([torch.nn.Linear(D_MODEL, 4) for _ in range(2)])
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected  # the reference's keys, one for one
    return m.to(device=device, dtype=dtype).eval()


def run(name, device, dtype, fused, tracing, graph=False):
    z, sd = load(name)
    m = build(name, sd, device, dtype, fused)
    t = lambda k: torch.from_numpy(z[k]).to(device=device, dtype=dtype)
    shapes, start = torch.from_numpy(z["shapes"]).to(device), torch.from_numpy(z["start"]).to(device)
    mask = torch.from_numpy(z["mask"]).to(device)
    kw = {"is_tracing": None} if tracing else {}
    if name.startswith("encoder"):
        src, pos = t("src").requires_grad_(not graph), t("pos").requires_grad_(not graph)
        if graph:
            g = T.GraphedModule(m, src, shapes, start, t("valid_ratios"), pos, mask, spatial_shapes_host=list(LEVELS))
            out = g(src, shapes, start, t("valid_ratios"), pos, mask).clone()
            return {"out": out.double().cpu().numpy()}
        out = m(src, shapes, start, t("valid_ratios"), pos, mask, **kw)
        out.backward(t("grad_out"))
        res = {"out": out, "g_src": src.grad, "g_pos": pos.grad}
    else:
        tgt, qp, mem = t("tgt").requires_grad_(True), t("query_pos").requires_grad_(True), t("memory").requires_grad_(True)
        r = m(tgt, t("reference_points"), mem, shapes, start, t("valid_ratios"), qp, mask, **kw)
        r["hs"].backward(t("grad_hs"))
        res = {"hs": r["hs"], "inter_references_out": r["inter_references_out"], "g_tgt": tgt.grad, "g_query_pos": qp.grad,
               "g_memory": mem.grad}
    return {k: v.detach().double().cpu().numpy() for k, v in res.items()}


def compare(got, name, rtol):
    z, _ = load(name)
    for k, v in got.items():
        want = z[k].astype(np.float64)
        assert_close(v, want, rtol, rtol * rms(want), f"{name}:{k}")


def test_fixtures_present():
    assert sorted(os.path.splitext(f)[0] for f in os.listdir(GOLD)) == sorted(CASES)


@pytest.mark.parametrize("name", CASES)
def test_layer_loop_mirror_matches_reference_on_the_tracing_branch(name):
    compare(run(name, "cpu", torch.float64, fused=False, tracing=True), name, 5e-6)  # fixtures are stored in fp32


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True], ids=["unfused", "fused"])
@pytest.mark.parametrize("name", CASES)
def test_layer_loop_on_gpu_matches_reference(name, fused, cuda_device):
    compare(run(name, cuda_device, torch.float32, fused=fused, tracing=False), name, 2e-4)


@pytest.mark.gpu
def test_graphed_encoder_replays_the_same_result(cuda_device):
    compare(run("encoder2", cuda_device, torch.float32, fused=True, tracing=False, graph=True), "encoder2", 2e-4)
