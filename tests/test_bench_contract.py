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


def _load_bench():
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_module_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_both_arms_print_the_same_workload_string():
    """The driver compares config.workload of the two arms: both come from ONE function."""
    bench = _load_bench()
    from aloception_oss_b200.synthetic import WORKLOADS

    s = bench.workload_string(WORKLOADS["C2"], "unit")
    assert s.startswith("C2: N=2 per GPU") and "Lq=300" in s and "M=8" in s and "P=4" in s and "D=32" in s
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "3")
    assert json.loads(r.stdout.strip().splitlines()[-1])["config"]["workload"] == s


def test_touched_bytes_counts_rows_where_taps_land():
    """roofline.touched_bytes: value rows are counted only where a bilinear tap lands (distinct (image, pixel, head) rows)."""
    import torch

    bench = _load_bench()
    from aloception_oss_b200.synthetic import Workload, level_tensors

    w = Workload("tb", 1, ((4, 4), (2, 2)), 1, M=2, P=1, D=8)
    shapes, start = level_tensors(w.levels)
    # head 0: level 0 at pixel centre (1, 1) -> one tap with weight 1 but the 2x2 window is (0..1, 0..1)?  centre of pixel (1,1)
    # is loc = 1.5/4 -> x = y = 1.0 exactly: floor = 1, taps (1,1), (1,2), (2,1), (2,2) all inside -> 4 rows;
    # level 1 at loc = -0.3 -> outside the (-1, size) window -> dropped, no rows
    # head 1: level 0 at loc 0 -> x = -0.5: taps x in {-1, 0}, y in {-1, 0}: only (0, 0) inside -> 1 row; level 1 centre of
    # pixel (0, 0) = 0.25 -> x = 0: taps (0,0), (0,1), (1,0), (1,1) -> 4 rows
    loc = torch.tensor([[[[[[0.375, 0.375]], [[-0.3, -0.3]]], [[[0.0, 0.0]], [[0.25, 0.25]]]]]])
    assert loc.shape == (1, 1, 2, 2, 1, 2)
    s = {"loc": loc, "shapes": torch.from_numpy(shapes), "start": torch.from_numpy(start)}
    tb = bench.touched_bytes(torch, w, s, 4)
    assert tb["value_rows_touched"] == 4 + 0 + 1 + 4 and tb["value_rows_total"] == 1 * 20 * 2
    nqmlp, nqmd, nsmd = 1 * 1 * 2 * 2 * 1, 1 * 1 * 2 * 8, 1 * 20 * 2 * 8
    assert tb["fwd"] == 4 * (9 * 8 + 3 * nqmlp + nqmd) + 12 * 2
    assert tb["bwd"] == 4 * (9 * 8 + nsmd + 6 * nqmlp + nqmd) + 12 * 2


def test_north_star_census_splits_the_global_batch_over_the_ranks():
    """north_star_census_b32: 32 images over `world` ranks, remainder to the first ranks; a rank without images reports nothing."""
    import inspect

    bench = _load_bench()
    src = inspect.getsource(bench.north_star_census)
    assert "global_batch // world + (1 if rank < global_batch % world else 0)" in src
    for world in (1, 2, 3, 4, 8, 40):
        per = [32 // world + (1 if r < 32 % world else 0) for r in range(world)]
        assert sum(per) == 32 and max(per) - min(per) <= 1
    assert bench.north_star_census(None, None, None, None, None, 40, 39) is None  # rank 39 of 40 has no image
