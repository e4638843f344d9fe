"""GPU tests of the two raw-pointer consumers of the forward kernel, called through the C ABI with ctypes only:

* ``msda_im2col_inference`` -- twin of the kernel wrapper behind the reference's TensorRT plugin
  (alonet/torch2trt/plugins/ms_deform_im2col/sources/ms_deform_im2col_kernel.cu:261-327), on the plugin test's own
  recipe (plugins/ms_deform_im2col/test.py:104-113: N, M, D = 1, 8, 32; Lq = 12000; levels 64^2 .. 8^2; L = P = 4);
* ``msda_forward_host`` -- the same operator on HOST buffers (stream-ordered staging inside the library).

The checker is the CPU oracle (oracle/msda_oracle.py); tolerances: fp32 rtol 1e-4, fp16 storage rtol 1e-2.
"""
import ctypes

import numpy as np
import pytest
import torch

from aloception_oss_b200 import _capi
from aloception_oss_b200.synthetic import Workload, host_inputs
from oracle import msda_oracle
from tests._util import assert_close, rms

pytestmark = pytest.mark.gpu

PLUGIN_TEST = Workload("trt_plugin_test", 1, ((64, 64), (32, 32), (16, 16), (8, 8)), 12000, M=8, P=4, D=32)
KFLOAT, KHALF, KINT8 = 0, 1, 2  # nvinfer1::DataType


def _oracle(x):
    return msda_oracle.forward(x["value"].astype(np.float64), x["shapes"], x["loc"].astype(np.float64),
                               x["attn"].astype(np.float64), x["start"])


@pytest.mark.parametrize("batch", [1, 3])
@pytest.mark.parametrize("data_type", [KFLOAT, KHALF], ids=["kFLOAT", "kHALF"])
def test_im2col_inference_on_the_plugin_test_recipe(data_type, batch, cuda_device):
    w = PLUGIN_TEST.with_batch(batch)
    if batch > 1:
        w = Workload(w.name, batch, w.levels, 1500, w.M, w.P, w.D)
    x = host_inputs(w, seed=5, loc_mode="wide")
    tdt = torch.float32 if data_type == KFLOAT else torch.float16
    dev = {k: torch.from_numpy(x[k]).to(cuda_device) for k in ("value", "loc", "attn", "shapes", "start")}
    for k in ("value", "loc", "attn"):
        dev[k] = dev[k].to(tdt).contiguous()
    out = torch.empty((w.N, w.Lq, w.M * w.D), dtype=tdt, device=cuda_device)
    stream = torch.cuda.current_stream(cuda_device).cuda_stream
    n0 = _capi.kernel_launch_count()
    rc = _capi.lib().msda_im2col_inference(
        ctypes.c_void_p(stream), dev["value"].data_ptr(), dev["shapes"].data_ptr(), dev["start"].data_ptr(),
        dev["loc"].data_ptr(), dev["attn"].data_ptr(), w.N, w.S, w.M, w.D, w.L, w.Lq, w.P, out.data_ptr(), data_type)
    assert rc == 0, _capi.last_error()
    torch.cuda.synchronize()
    assert _capi.kernel_launch_count() == n0 + 1
    # the oracle sees exactly the values the kernel saw (fp16-rounded for kHALF)
    xr = dict(x, value=dev["value"].float().cpu().numpy(), loc=dev["loc"].float().cpu().numpy(),
              attn=dev["attn"].float().cpu().numpy())
    want = _oracle(xr)
    got = out.double().cpu().numpy()
    if data_type == KFLOAT:
        assert_close(got, want, 1e-4, 1e-7 * rms(want), "data_col")
    else:
        assert_close(got, want, 1e-2, 1e-2 * rms(want), "data_col")


def test_im2col_inference_rejects_other_trt_types(cuda_device):
    rc = _capi.lib().msda_im2col_inference(None, None, None, None, None, None, 1, 4, 1, 32, 1, 1, 1, None, KINT8)
    assert rc == -1 and "unsupported nvinfer1::DataType" in _capi.last_error()


@pytest.mark.parametrize("pinned", [True, False], ids=["pinned", "pageable"])
def test_forward_host_buffers(pinned, cuda_device):
    w = Workload("host", 2, ((20, 27), (10, 14), (5, 7), (3, 4)), 70, M=8, P=4, D=32)
    x = host_inputs(w, seed=9, loc_mode="wide")
    t = {k: torch.from_numpy(x[k]) for k in ("value", "loc", "attn", "shapes", "start")}
    out = torch.empty((w.N, w.Lq, w.M * w.D), dtype=torch.float32)
    if pinned:
        t = {k: v.pin_memory() for k, v in t.items()}
        out = out.pin_memory()
    dims = _capi.MsdaDims(w.N, w.S, w.M, w.D, w.L, w.Lq, w.P)
    with torch.cuda.device(cuda_device):
        stream = torch.cuda.current_stream().cuda_stream
        rc = _capi.lib().msda_forward_host(t["value"].data_ptr(), t["shapes"].data_ptr(), t["start"].data_ptr(),
                                           t["loc"].data_ptr(), t["attn"].data_ptr(), out.data_ptr(),
                                           ctypes.byref(dims), _capi.F32, ctypes.c_void_p(stream))
    assert rc == 0, _capi.last_error()  # the call returns after `out` is complete on the host
    want = _oracle(x)
    assert_close(out.double().numpy(), want, 1e-4, 1e-7 * rms(want), "out")
