"""Shared helpers for the test-suite: golden fixture loading and tolerance checks."""
from __future__ import annotations

import glob
import os

import numpy as np

from aloception_oss_b200.synthetic import Workload, host_inputs

GOLD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLD_DIR, "*.npz")))


def load_golden(name):
    """Returns (workload, numpy inputs in the reference dtype, dict of reference outputs)."""
    z = np.load(os.path.join(GOLD_DIR, name + ".npz"))
    w = Workload(
        str(z["name"]), int(z["N"]), tuple((int(h), int(x)) for h, x in z["levels"]), int(z["Lq"]),
        M=int(z["M"]), P=int(z["P"]), D=int(z["D"]),
    )
    dt = np.float64 if str(z["ref_dtype"]) == "f64" else np.float32
    x = host_inputs(w, int(z["seed"]), str(z["loc_mode"]), dtype=dt)
    ref = {k: z[k] for k in z.files}
    return w, x, ref


def assert_close(got, want, rtol, atol, what=""):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    err = np.abs(got - want)
    tol = atol + rtol * np.abs(want)
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(
            f"{what}: {bad.sum()} / {bad.size} elements out of tolerance (rtol={rtol}, atol={atol}); "
            f"worst at {i}: got {got[i]!r}, want {want[i]!r}, |err| {err[i]:.3e}"
        )


def rms(a):
    """Root-mean-square over the NON-ZERO entries (grad_value is mostly structural zeros)."""
    a = np.asarray(a, dtype=np.float64)
    a = a[a != 0]
    return float(np.sqrt((a * a).mean())) if a.size else 0.0


def assert_close_grad(got, want, rtol=1e-4, what=""):
    """Gradient tolerance for fp32 work: ``rtol`` relative + ``rtol * rms(want)`` absolute.

    grad_loc / grad_attn are channel sums with cancellation (differences of neighbouring taps), so a
    purely relative bound cannot hold for ANY fp32 evaluation order -- the reference's own fp32 run
    differs from its fp64 run by up to 6e-5 * rms on these fixtures.  SURVEY.md section 7, hard part 5.
    """
    assert_close(got, want, rtol, rtol * rms(want), what)


def check_grad_value(gv, ref, rtol, atol=None, what="grad_value"):
    """grad_value against a golden record (full tensor, or projections + picked entries).

    ``atol=None`` -> ``rtol * rms`` of the reference entries: a bilinear weight is ``1 - frac(y)`` with
    ``y ~ 100`` known to ~4e-6 in fp32, so small weights carry a large RELATIVE error in any fp32 code.
    """
    gv = np.asarray(gv, dtype=np.float64)
    if atol is None:
        atol = rtol * rms(ref["grad_value"] if "grad_value" in ref else ref["gv_pick_val"])
    if "grad_value" in ref:
        assert_close(gv, ref["grad_value"], rtol, atol, what)
        return
    D = gv.shape[-1]
    S = gv.shape[1]
    assert_close(gv.sum(-1), ref["gv_sum_channels"], rtol, atol * D, what + ".sum(channels)")
    assert_close(gv.sum(1), ref["gv_sum_pixels"], rtol, atol * np.sqrt(S) * 4, what + ".sum(pixels)")
    assert_close(gv.reshape(-1)[ref["gv_pick_idx"]], ref["gv_pick_val"], rtol, atol, what + "[picked]")


def near_floor_discontinuity(loc, shapes, eps=2e-5):
    """Mask (N, Lq, M, L, P) of samples whose pixel coordinate is within ``eps`` of an integer.

    grad_loc is DISCONTINUOUS there (floor() switches the interpolated pixel pair), so two correct evaluations
    that round ``loc * size - 0.5`` differently (fp32 FMA vs fp32 mul+sub vs fp64) legitimately disagree on
    grad_loc for such a sample -- the reference's own CUDA fp32 kernel and its fp64 evaluation do
    (tests/debug_gradloc.py).  out, grad_value and grad_attn are continuous and are never masked.
    """
    loc = np.asarray(loc, dtype=np.float64)
    shapes = np.asarray(shapes, dtype=np.float64)
    wh = shapes[:, ::-1].reshape(1, 1, 1, -1, 1, 2)  # (W, H) per level, matching (x, y)
    pix = loc * wh - 0.5
    return (np.abs(pix - np.round(pix)) < eps).any(-1)


def assert_close_grad_loc(got, want, loc, shapes, rtol=1e-4, what="grad_loc"):
    """grad_loc check that skips the (measure-zero) samples sitting on a floor() discontinuity."""
    got = np.asarray(got, dtype=np.float64).copy()
    want = np.asarray(want, dtype=np.float64)
    skip = near_floor_discontinuity(loc, shapes)
    assert skip.mean() < 1e-2, "too many samples masked"
    got[skip] = want[skip]
    assert_close(got, want, rtol, rtol * rms(want), what)
