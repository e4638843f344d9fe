"""ctypes front-end of the C oracle (oracle/msda_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of msda_oracle.c.  Imported by tests/,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of bench.py,
never by the product package ``aloception_oss_b200``.

The oracle restates the reference arithmetic
(alonet/deformable_detr/ops/src/cuda/ms_deform_im2col_cuda.cuh:33-159,237-299) on the CPU.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "msda_oracle.c")
_SO = os.path.join(_HERE, "libmsda_oracle.so")
_lock = threading.Lock()
_lib = None


def build(force: bool = False) -> str:
    """Compile msda_oracle.c with gcc (OpenMP) next to its source; returns the .so path."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-o", _SO, _SRC, "-lm"]
        subprocess.run(cmd, check=True, capture_output=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    with _lock:
        if _lib is None:
            _lib = ctypes.CDLL(build())
            i = ctypes.c_int
            vp = ctypes.c_void_p
            for suf in ("f32", "f64"):
                getattr(_lib, f"msda_oracle_forward_{suf}").argtypes = [vp] * 6 + [i] * 7
                getattr(_lib, f"msda_oracle_forward_{suf}").restype = None
                getattr(_lib, f"msda_oracle_backward_{suf}").argtypes = [vp] * 9 + [i] * 7
                getattr(_lib, f"msda_oracle_backward_{suf}").restype = None
    return _lib


def _prep(value, shapes, start, loc, attn):
    value = np.ascontiguousarray(value)
    if value.dtype not in (np.float32, np.float64):
        raise TypeError("oracle works in float32 or float64")
    dt = value.dtype
    loc = np.ascontiguousarray(loc, dtype=dt)
    attn = np.ascontiguousarray(attn, dtype=dt)
    shapes = np.ascontiguousarray(shapes, dtype=np.int32)
    if start is None:
        hw = shapes[:, 0].astype(np.int64) * shapes[:, 1]
        start = np.concatenate([[0], np.cumsum(hw)[:-1]])
    start = np.ascontiguousarray(start, dtype=np.int32)
    N, S, M, D = value.shape
    _, Lq, M2, L, P, two = loc.shape
    assert M2 == M and two == 2 and attn.shape == (N, Lq, M, L, P) and shapes.shape == (L, 2)
    return value, shapes, start, loc, attn, (N, S, M, D, L, Lq, P), ("f32" if dt == np.float32 else "f64")


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def forward(value, shapes, loc, attn, start=None) -> np.ndarray:
    """out (N, Lq, M*D) for numpy inputs; dtype follows ``value`` (float32 or float64)."""
    value, shapes, start, loc, attn, dims, suf = _prep(value, shapes, start, loc, attn)
    N, S, M, D, L, Lq, P = dims
    out = np.empty((N, Lq, M * D), dtype=value.dtype)
    getattr(lib(), f"msda_oracle_forward_{suf}")(_p(value), _p(shapes), _p(start), _p(loc), _p(attn), _p(out), *dims)
    return out


def backward(grad_out, value, shapes, loc, attn, start=None):
    """(grad_value, grad_loc, grad_attn) for numpy inputs."""
    value, shapes, start, loc, attn, dims, suf = _prep(value, shapes, start, loc, attn)
    N, S, M, D, L, Lq, P = dims
    grad_out = np.ascontiguousarray(grad_out, dtype=value.dtype).reshape(N, Lq, M * D)
    gv = np.empty_like(value)
    gl = np.empty_like(loc)
    ga = np.empty_like(attn)
    getattr(lib(), f"msda_oracle_backward_{suf}")(
        _p(grad_out), _p(value), _p(shapes), _p(start), _p(loc), _p(attn), _p(gv), _p(gl), _p(ga), *dims
    )
    return gv, gl, ga


# ---- torch conveniences (CPU tensors in, CPU tensors out) --------------------------------------
def forward_t(value, shapes, loc, attn, start=None, as_double=False):
    import torch

    f = (lambda t: t.detach().double().cpu().numpy()) if as_double else (lambda t: t.detach().cpu().numpy())
    out = forward(f(value), shapes.cpu().numpy(), f(loc), f(attn), None if start is None else start.cpu().numpy())
    return torch.from_numpy(out)


def backward_t(grad_out, value, shapes, loc, attn, start=None, as_double=False):
    import torch

    f = (lambda t: t.detach().double().cpu().numpy()) if as_double else (lambda t: t.detach().cpu().numpy())
    gs = backward(
        f(grad_out), f(value), shapes.cpu().numpy(), f(loc), f(attn), None if start is None else start.cpu().numpy()
    )
    return tuple(torch.from_numpy(g) for g in gs)
