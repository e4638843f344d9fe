"""bench.py contract pieces that run without a GPU: the reference arm (`--impl reference` times the reference's CPU path and
prints exactly ONE JSON line on stdout with the agreed keys), and the helpers that keep stdout clean."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ, **(env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = run_bench("--impl", "reference", "--steps", "2", "--warmup", "3", "--gpus", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gsamples/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] >= 3 and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "C2" in d["config"]["workload"] and "model" not in d["config"]
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = run_bench("--impl", "reference", "--steps", "1", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_own_arm_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        return
    r = run_bench("--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-e2e")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
