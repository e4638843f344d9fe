"""The C ABI from a plain-C host (tests/c_abi/c_abi_smoke.c): no Python, no torch on the data path.

CPU: the header is valid C11 and every entry point the program uses links against libmsda_b200.so.
GPU: the program runs forward / backward / host-buffer forward / the TensorRT-plugin twin on cudaMalloc'ed tensors and
checks them against the C oracle.
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "c_abi_smoke.c")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build(tmp_path):
    from aloception_oss_b200 import _capi
    from oracle import msda_oracle

    _capi.build_library()
    msda_oracle.build()
    exe = str(tmp_path / "c_abi_smoke")
    pkg, ora = os.path.join(ROOT, "aloception_oss_b200"), os.path.join(ROOT, "oracle")
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CUDA, "include"), SRC,
           "-o", exe, "-L" + pkg, "-lmsda_b200", "-L" + ora, "-lmsda_oracle", "-L" + os.path.join(CUDA, "lib64"), "-lcudart", "-lm",
           "-Wl,-rpath," + pkg, "-Wl,-rpath," + ora, "-Wl,-rpath," + os.path.join(CUDA, "lib64")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_plain_c_program_compiles_and_links(tmp_path):
    assert os.path.exists(build(tmp_path))


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_plain_c_program_matches_the_oracle(tmp_path, cuda_device):
    exe = build(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "C ABI smoke: OK" in res.stdout
    for name in ("forward", "grad_value", "grad_attn", "grad_loc", "forward_host", "plugin_twin"):
        assert f"{name:<12s} ok" in res.stdout, res.stdout
    # deterministic backward (bit-reproducible), msda_forward_ws with the opt-in SM-affine schedule, bf16 value + fp32 loc / attn
    assert "grad_value (deterministic)" in res.stdout and "forward_ws (paired, SM-affine)" in res.stdout, res.stdout
    assert "forward (bf16 value, fp32 loc / attn): max |err|" in res.stdout, res.stdout


# ------------------------------------------------------------------ include/sortv_b200.h

SORTV_SRC = os.path.join(ROOT, "tests", "c_abi", "sortv_smoke.c")


def build_sortv(tmp_path):
    from aloception_oss_b200 import rotated_iou
    from oracle import sortv_oracle

    rotated_iou.build_library()
    sortv_oracle.build()
    exe = str(tmp_path / "sortv_smoke")
    pkg, ora = os.path.join(ROOT, "aloception_oss_b200"), os.path.join(ROOT, "oracle")
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CUDA, "include"),
           SORTV_SRC, "-o", exe, "-L" + pkg, "-lsortv_b200", "-L" + ora, "-lsortv_oracle", "-L" + os.path.join(CUDA, "lib64"),
           "-lcudart", "-lm", "-Wl,-rpath," + pkg, "-Wl,-rpath," + ora, "-Wl,-rpath," + os.path.join(CUDA, "lib64")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_sortv_plain_c_program_compiles_and_links(tmp_path):
    assert os.path.exists(build_sortv(tmp_path))


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_sortv_plain_c_program_matches_the_oracle(tmp_path, cuda_device):
    exe = build_sortv(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "sortv C ABI smoke: OK" in res.stdout
    for v in range(5):
        assert f"variant {v}     ok" in res.stdout, res.stdout


# ------------------------------------------------------------------ TensorRT plugin class

PLUGIN_SRC = os.path.join(ROOT, "aloception_oss_b200", "csrc", "trt_plugin", "msda_trt_plugin.cpp")
PLUGIN_TEST = os.path.join(ROOT, "tests", "c_abi", "trt_plugin_smoke.cpp")


def build_plugin_smoke(tmp_path):
    """The plugin source + its driver, compiled against the MOCK TensorRT header (TensorRT is not in this image)."""
    from aloception_oss_b200 import _capi

    _capi.build_library()
    exe = str(tmp_path / "trt_plugin_smoke")
    pkg = os.path.join(ROOT, "aloception_oss_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wno-comment", "-I" + os.path.join(ROOT, "tests", "c_abi", "mock_tensorrt"),
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CUDA, "include"), PLUGIN_SRC, PLUGIN_TEST, "-o", exe,
           "-L" + pkg, "-lmsda_b200", "-L" + os.path.join(CUDA, "lib64"), "-lcudart",
           "-Wl,-rpath," + pkg, "-Wl,-rpath," + os.path.join(CUDA, "lib64")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_trt_plugin_class_compiles_against_the_mock_interface(tmp_path):
    assert os.path.exists(build_plugin_smoke(tmp_path))


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_trt_plugin_class_enqueue_matches_the_operator(tmp_path, cuda_device):
    exe = build_plugin_smoke(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "TRT plugin smoke: OK" in res.stdout and "enqueue      ok" in res.stdout
