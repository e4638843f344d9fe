#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 multi-scale deformable attention operator.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C2] [--dtype f32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one forward + one backward of the operator over one batch of synthetic input
(BASELINE.json metric: "MSDeformAttn fwd+bwd Gsamples/s at COCO 4-level shapes"; sample = one (n, q, m, l, p)
sampling point).  Default workload = BASELINE.json configs[1]/[2] ("C2": N=2 per GPU, levels 100/50/25/13 squared,
300 queries, 8 heads, 4 points, d=256).  Multi-GPU = weak scaling: every rank owns its own batch shard, no
collective on the data path; one NCCL broadcast of the level metadata at set-up.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how every field is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "MSDeformAttn fwd+bwd Gsamples/s (COCO 4-level shapes)"
UNIT = "Gsamples/s"
L2_BYTES = 126 * 1024 * 1024
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16", "f16"])
    ap.add_argument("--loc-mode", default="unit")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-bind", action="store_true", help="do not pin the process to the GPU's NUMA-local CPUs")
    ap.add_argument("--extra", action="store_true", help="also time the other COCO shapes (reported under 'extra')")
    ap.add_argument("--no-north-star", action="store_true", help="skip the forward-only timing of the 800x1333 / 300-query shapes")
    return ap.parse_args()


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU through NVML every 20 ms while running."""

    BAD = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80}
    NOTE = {"sw_power_cap": 0x4, "sync_boost": 0x10, "display_clocks": 0x100}

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        med = int(statistics.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU implementation of the path (pure PyTorch), ported
# ------------------------------------------------------------------------------------------------------------
def workload_string(w, loc_mode):
    """config.workload: ONE string for both arms (the driver compares them)."""
    return (f"{w.name}: N={w.N} per GPU, levels={[list(l) for l in w.levels]}, Lq={w.Lq}, M={w.M}, P={w.P}, D={w.D}, "
            f"loc={loc_mode}; step = forward + backward")


def cpu_reference_fn():
    """The CPU implementation of the path that is timed as the baseline: the reference's OWN
    ``ms_deform_attn_core_pytorch`` + autograd (unmodified file bundled into the git-ignored oracle/_ref/ by
    oracle/build_ref_py.py, kind "reference"); the port oracle/msda_torch_port.py (kind "port") only when that bundle is
    absent.  Returns (callable, kind, description)."""
    try:
        from oracle import build_ref_py

        if build_ref_py.bundled():
            build_ref_py.load()
            return build_ref_py.fwd_bwd, "reference", ("ms_deform_attn_core_pytorch + autograd, unmodified "
                                                       "alonet/deformable_detr/ops/functions/ms_deform_attn_func.py:85-190 (oracle/_ref/alonet_ref_py)")
    except Exception:
        pass
    from oracle.msda_torch_port import msda_fwd_bwd_port

    return msda_fwd_bwd_port, "port", "oracle/msda_torch_port.py (port of ms_deform_attn_func.py:85-190) + autograd"


def cpu_port_rate(workload, budget_s, steps=None, loc_mode="unit", warmup=1):
    """Times forward + autograd backward of the reference's CPU path on the host cores.

    Returns (Gsamples/s, ms per step, steps run, sample description, threads, kind)."""
    import torch

    from aloception_oss_b200.synthetic import Workload, torch_inputs

    fn, kind, what = cpu_reference_fn()
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    w = workload
    x = torch_inputs(w, seed=3, loc_mode=loc_mode)
    t0 = time.perf_counter()
    fn(x["value"], x["shapes"], x["loc"], x["attn"], x["grad_out"])  # warm-up + cost probe
    probe = time.perf_counter() - t0
    t0 = time.perf_counter()
    fn(x["value"], x["shapes"], x["loc"], x["attn"], x["grad_out"])  # second probe: the first pays one-off allocations
    probe = min(probe, time.perf_counter() - t0)
    desc = f"{w.name}: full batch N={w.N}, Lq={w.Lq}"
    if steps is not None and steps * probe > budget_s and w.Lq > 1:
        # bound the run: keep the full value pyramid, take a prefix of the queries
        lq = max(1, int(w.Lq * budget_s / (steps * probe)))
        w = Workload(w.name, w.N, w.levels, lq, w.M, w.P, w.D)
        x = dict(x, loc=x["loc"][:, :lq].contiguous(), attn=x["attn"][:, :lq].contiguous(),
                 grad_out=x["grad_out"][:, :lq].contiguous())
        desc = f"{workload.name}: N={w.N}, first {lq} of {workload.Lq} queries per image (bounded sample)"
    n = steps if steps is not None else max(3, min(200, int(budget_s / max(probe, 1e-4))))
    for _ in range(max(0, warmup - 2)):  # the two cost probes above were the first warm-up steps
        fn(x["value"], x["shapes"], x["loc"], x["attn"], x["grad_out"])
    t0 = time.perf_counter()
    for _ in range(n):
        fn(x["value"], x["shapes"], x["loc"], x["attn"], x["grad_out"])
    dt = time.perf_counter() - t0
    return w.samples * n / dt / 1e9, dt / n * 1e3, n, f"{desc}; {what}", threads, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from aloception_oss_b200.synthetic import WORKLOADS

    w = WORKLOADS[args.workload]
    warm = max(3, min(args.warmup, 10))  # W >= 3 untimed steps (bounded: a CPU step takes ~20 ms)
    rate, ms, n, desc, threads, kind = cpu_port_rate(w, budget_s=120.0, steps=args.steps, loc_mode=args.loc_mode, warmup=warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(w, args.loc_mode),
                   "implementation": "the reference's pure-PyTorch CPU path on the host cores (rank 0 only)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------------------
_JSON_FD = None


def quiet_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL prints its version banner on fd 1) go to stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def bind_near_gpu(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index` (NUMA node of its PCIe root), BEFORE any pinned
    host buffer is allocated, so that the e2e staging memory is first-touched next to the GPU.  Returns a short note."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = len(os.sched_getaffinity(0))
        return f"cpu affinity {before} -> {after} cpus (nvmlDeviceSetCpuAffinity)"
    except Exception as e:  # no NUMA information in this VM, or not permitted: run unpinned
        return f"unpinned ({type(e).__name__})"


def main():
    args = parse_args()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import aloception_oss_b200 as msda
    from aloception_oss_b200 import _capi
    from aloception_oss_b200.synthetic import WORKLOADS, device_inputs, level_tensors

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    all_cpus = os.sched_getaffinity(0)
    numa_note = bind_near_gpu(local) if not args.no_bind else "unpinned (--no-bind)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    msda.load_ops()

    tdt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[args.dtype]
    elt = 4 if args.dtype == "f32" else 2
    w = WORKLOADS[args.workload]  # per-GPU shard (weak scaling)

    # the only replicated state: level metadata, broadcast once from rank 0 (no collective on the data path)
    shapes_np, start_np = level_tensors(w.levels)
    meta = torch.from_numpy(shapes_np).to(dev).reshape(-1) if rank == 0 else torch.zeros(2 * w.L, dtype=torch.int32, device=dev)
    if world > 1:
        dist.broadcast(meta, src=0)
    assert meta.cpu().tolist() == shapes_np.reshape(-1).tolist()

    # rotating input sets: total footprint > 2x L2 so that every step reads its inputs from HBM
    step_bytes = w.algorithmic_bytes(elt, False) + w.algorithmic_bytes(elt, True)
    # (the gather touches only part of each value tensor, so provision ~8x L2 of nominal footprint)
    n_sets = max(4, int(8 * L2_BYTES / step_bytes) + 2)
    while n_sets * step_bytes > 40e9 and n_sets > 2:
        n_sets -= 1
    sets = [device_inputs(w, seed=1000 * rank + i, device=dev, dtype=tdt, loc_mode=args.loc_mode) for i in range(n_sets)]
    # result buffers ROTATE with the input sets (a freshly re-allocated grad_value would sit at the same address every
    # step and stay L2-resident: 13.1 vs 14.5 us for the C2 backward, profiles/README.md)
    for s_ in sets:
        s_["out"] = torch.empty((w.N, w.Lq, w.M * w.D), dtype=tdt, device=dev)
        s_["grads"] = [torch.empty_like(s_["value"]), torch.empty_like(s_["loc"]), torch.empty_like(s_["attn"])]

    def fwd(s):
        return msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], out=s["out"])

    def bwd(s):
        return msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"],
                                            grads=s["grads"])

    def step(i):
        s = sets[i % n_sets]
        return fwd(s), bwd(s)

    K, W = args.steps, max(args.warmup, 3)
    for i in range(W):
        step(i)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def capture(fn, n):
        """CUDA graph of n consecutive calls fn(i); returns (graph, launches recorded)."""
        g = torch.cuda.CUDAGraph()
        c0 = _capi.kernel_launch_count()
        with torch.cuda.graph(g):
            for i in range(n):
                fn(i)
        return g, _capi.kernel_launch_count() - c0

    def time_graph(g, replays=1, repeats=1):
        """`repeats` measurements of `replays` back-to-back graph replays each (events on the launching stream, barrier +
        synchronize on both sides, max over ranks per measurement).  Returns the list of measured milliseconds."""
        g.replay()  # untimed: graph upload / first-replay cost
        out = []
        for _ in range(repeats):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(replays):
                g.replay()
            e1.record()
            barrier()
            out.append(e0.elapsed_time(e1))
        if world > 1:
            t = torch.tensor(out, device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out = [float(v) for v in t.tolist()]
        return out

    clocks = ClockSampler(local).start()

    # ---- headline: K steps (fwd + bwd), one CUDA graph, device-resident inputs ---------------------------------
    chunk = min(K, 500)  # graph of `chunk` steps replayed K/chunk times (K rounded down to a multiple)
    reps = max(1, K // chunk)
    K_eff = chunk * reps
    # One measurement = exactly K_eff steps.  A single short measurement is jitter-limited (K = 20 is 0.4 ms), so the
    # measurement is repeated and the MEDIAN per-step time is reported (`steps` stays K_eff; all repeats under "timing").
    repeats = max(5, min(50, -(-25000 // K_eff)))
    g_step, launches = capture(step, chunk)
    ms_all = time_graph(g_step, reps, repeats)
    ms_per_step = statistics.median(ms_all) / K_eff
    value = w.samples * world / (ms_per_step * 1e-3) / 1e9

    # ---- per-pass timing for the roofline: forward-only and backward-only graphs -----------------------------
    g_f, lf = capture(lambda i: fwd(sets[i % n_sets]), chunk)
    ms_fwd = statistics.median(time_graph(g_f, reps, repeats)) / K_eff
    g_b, lb = capture(lambda i: bwd(sets[i % n_sets]), chunk)
    ms_bwd = statistics.median(time_graph(g_b, reps, repeats)) / K_eff
    peak, peak_src = hbm_peak()
    touched = touched_bytes(torch, w, sets[0], elt)

    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(f"{w.name}/{args.dtype}", {})
    except Exception:
        traffic = {}

    def roof(bytes_, ms, what, tkey=None):
        ach = bytes_ / (ms * 1e-3) / 1e9
        tb = touched[tkey]
        return {"bound": "hbm", "kernel": what, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                "frac": round(ach / peak, 4), "traffic": traffic.get(tkey), "traffic_source": "ncu --set full, profiles/traffic.json (cold-cache single launch)",
                "algorithmic_bytes": bytes_, "us_per_launch": round(ms * 1e3, 3),
                # SURVEY 8(d) counts the whole value tensor once although a 300-query call touches part of it: `frac` can
                # exceed 1 on large decoder shapes.  touched_bytes counts value / grad_value rows only where a tap lands
                # (computed from this run's sampling locations, 128-byte (row, head) granules) -- the bytes that must move.
                "touched_bytes": tb, "frac_touched": round(tb / (ms * 1e-3) / 1e9 / peak, 4),
                "peak_source": peak_src}

    bwd_kernel = "msda_zero_kernel + msda_bwd_tile_kernel" if (w.Lq == w.S and args.dtype == "f32") else "msda_zero_kernel + msda_bwd_sg_kernel"
    r_fwd = roof(w.algorithmic_bytes(elt, False), ms_fwd, "msda_fwd_sg_kernel (forward pass)", "fwd")
    r_bwd = roof(w.algorithmic_bytes(elt, True), ms_bwd, bwd_kernel + " (backward pass, PDL-overlapped)", "bwd")
    dominant = r_bwd if ms_bwd >= ms_fwd else r_fwd

    # ---- eager (no graph) rate: what a Python caller gets launch-by-launch ------------------------------------
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_eager = min(K_eff, 300)
    e0.record()
    for i in range(n_eager):
        step(i)
    e1.record()
    barrier()
    ms_eager = e0.elapsed_time(e1) / n_eager

    # ---- end to end: pinned host inputs -> H2D -> fwd + bwd -> D2H of out + 3 grads, pipelined over 3 streams ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(torch, msda, sets, w, dev, world, dist, min(K_eff, 200), elt)

    clk = clocks.stop()

    extra = {}
    if args.extra and rank == 0:
        extra = run_extra(torch, msda, _capi, dev, tdt, elt, peak)
    # BASELINE.json's target is quoted on the 800x1333 pyramid with 300 queries: its forward is timed in every run (rank 0)
    north_star = north_star_forward(torch, msda, dev, tdt, elt, peak) if rank == 0 and not args.no_north_star else None
    # ... and configs[3] itself, B = 32 images batch-sharded over the ranks (strong scaling): the operator census of one
    # DeformableDETR-R50 inference forward, every rank takes part (max over ranks)
    census = north_star_census(torch, msda, dist, dev, tdt, world, rank) if not args.no_north_star else None
    census_train = training_census(torch, msda, dist, dev, tdt, world) if not args.no_north_star else None

    cpu_base = None
    if rank == 0 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)  # the CPU baseline gets every host core again
        rate, ms, n, desc, threads, kind = cpu_port_rate(w, budget_s=12.0, loc_mode=args.loc_mode)
        cpu_base = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
                    "sample": f"{desc}; {n} steps of fwd+autograd-bwd, {ms:.1f} ms/step"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K_eff, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {
                "workload": workload_string(w, args.loc_mode),
                "samples_per_step_per_gpu": w.samples,
                "l2_policy": f"rotating {n_sets} distinct sets of inputs AND result buffers (out, grad_value, grad_loc, grad_attn) "
                             f"({n_sets * step_bytes / 1e6:.0f} MB nominal, >= 8x the 126 MB L2): no tensor of a step is L2-resident from the step before",
                "launch": f"CUDA graph of {chunk} steps x {reps} replays = {K_eff} steps per measurement; {repeats} measurements, median reported",
                "sharding": "batch-sharded, no data-path collective",
                "host": numa_note,
            },
            "timing": {"measurements": repeats, "steps_per_measurement": K_eff,
                       "ms_per_step_min": min(ms_all) / K_eff, "ms_per_step_median": ms_per_step, "ms_per_step_max": max(ms_all) / K_eff},
            "roofline": dominant, "roofline_fwd": r_fwd, "roofline_bwd": r_bwd,
            "fwd_only": {"value": w.samples * world / (ms_fwd * 1e-3) / 1e9, "unit": UNIT, "us": ms_fwd * 1e3},
            "bwd_only": {"value": w.samples * world / (ms_bwd * 1e-3) / 1e9, "unit": UNIT, "us": ms_bwd * 1e3},
            "eager": {"value": w.samples * world / (ms_eager * 1e-3) / 1e9, "unit": UNIT, "us_per_step": ms_eager * 1e3},
            "cpu_baseline": cpu_base, "e2e": e2e,
            "gpu_launches": int(launches * reps), "launches_per_step": launches / chunk,
            "clocks": clk,
        }
        if north_star:
            line["north_star_forward"] = north_star
        if census:
            line["north_star_census_b32"] = census
        if census_train:
            line["training_census_b2_per_gpu"] = census_train
        if extra:
            line["extra"] = extra
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def touched_bytes(torch, w, s, elt):
    """Bytes a call MUST move given its sampling locations: like SURVEY 8(d)'s algorithmic bytes, but the value tensor
    (forward, backward gather) is counted only where a bilinear tap lands -- distinct (image, pixel, head) rows of D
    elements.  grad_value is still written in full (zero-fill).  Computed on the device from one input set."""
    N, S, M, D, L, P, Lq = w.N, w.S, w.M, w.D, w.L, w.P, w.Lq
    loc = s["loc"].float()
    shapes = s["shapes"].long()
    start = s["start"].long()
    hit = torch.zeros((N * S * M,), dtype=torch.bool, device=loc.device)
    n_idx = torch.arange(N, device=loc.device).view(N, 1, 1, 1)
    m_idx = torch.arange(M, device=loc.device).view(1, 1, M, 1)
    for l in range(L):
        H, W = int(shapes[l, 0]), int(shapes[l, 1])
        x = loc[:, :, :, l, :, 0] * W - 0.5
        y = loc[:, :, :, l, :, 1] * H - 0.5
        inside = (x > -1) & (y > -1) & (x < W) & (y < H)
        x0, y0 = torch.floor(x).long(), torch.floor(y).long()
        for dy in (0, 1):
            for dx in (0, 1):
                xx, yy = x0 + dx, y0 + dy
                ok = inside & (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
                row = (n_idx * S + int(start[l]) + yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)) * M + m_idx
                hit[row[ok]] = True
    rows = int(hit.sum().item())
    nsmd, nqmlp, nqmd = N * S * M * D, N * Lq * M * L * P, N * Lq * M * D
    return {"fwd": elt * (rows * D + 3 * nqmlp + nqmd) + 12 * L,
            "bwd": elt * (rows * D + nsmd + 6 * nqmlp + nqmd) + 12 * L,
            "value_rows_touched": rows, "value_rows_total": N * S * M}


def run_e2e(torch, msda, sets, w, dev, world, dist, K, elt):
    """Same step through the public API with HOST buffers: every step uploads its inputs from pinned memory and
    downloads out + the three gradients.  Inputs and results live in packed arenas (one H2D and one D2H copy per
    step); 3-slot pipeline on upload / compute / download streams."""
    names = ("value", "loc", "attn", "grad_out")
    dt = sets[0]["value"].dtype
    esz = sets[0]["value"].element_size()

    def layout(shapes):
        offs, off = [], 0
        for shp in shapes:
            n = 1
            for d in shp:
                n *= d
            offs.append((off, n, shp))
            off += (n * esz + 255) // 256 * 256 // esz
        return offs, off

    in_shapes = [tuple(sets[0][k].shape) for k in names]
    out_shapes = [(w.N, w.Lq, w.M * w.D), (w.N, w.S, w.M, w.D), (w.N, w.Lq, w.M, w.L, w.P, 2), (w.N, w.Lq, w.M, w.L, w.P)]
    in_lay, in_total = layout(in_shapes)
    out_lay, out_total = layout(out_shapes)

    def views(arena, lay):
        return [arena[o:o + n].view(shp) for o, n, shp in lay]

    n_host = min(len(sets), 4)
    host_in = []
    for i in range(n_host):
        arena = torch.empty(in_total, dtype=dt).pin_memory()
        for v, k in zip(views(arena, in_lay), names):
            v.copy_(sets[i][k])
        host_in.append(arena)
    depth = 3
    slots = []
    for _ in range(depth):
        d_in = torch.empty(in_total, dtype=dt, device=dev)
        d_out = torch.empty(out_total, dtype=dt, device=dev)
        slots.append({"d_in": d_in, "in": dict(zip(names, views(d_in, in_lay))), "d_out": d_out, "out": views(d_out, out_lay),
                      "h_out": torch.empty(out_total, dtype=dt).pin_memory(),
                      "ev_up": torch.cuda.Event(), "ev_done": torch.cuda.Event(), "ev_down": torch.cuda.Event()})
    shapes, start = sets[0]["shapes"], sets[0]["start"]
    s_up, s_cp, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    h2d = sum(n for _, n, _ in in_lay) * esz
    d2h = sum(n for _, n, _ in out_lay) * esz

    def run(n, compute=True):
        for i in range(n):
            sl = slots[i % depth]
            with torch.cuda.stream(s_up):
                s_up.wait_event(sl["ev_done"])  # slot inputs are free once the previous compute on it finished
                sl["d_in"].copy_(host_in[i % n_host], non_blocking=True)
                sl["ev_up"].record(s_up)
            with torch.cuda.stream(s_cp):
                s_cp.wait_event(sl["ev_up"])
                s_cp.wait_event(sl["ev_down"])  # slot outputs are free once their previous download finished
                x, o = sl["in"], sl["out"]
                if compute:
                    msda.ms_deform_attn_forward(x["value"], shapes, start, x["loc"], x["attn"], out=o[0])
                    msda.ms_deform_attn_backward(x["value"], shapes, start, x["loc"], x["attn"], x["grad_out"], grads=o[1:])
                sl["ev_done"].record(s_cp)
            with torch.cuda.stream(s_dn):
                s_dn.wait_event(sl["ev_done"])
                sl["h_out"].copy_(sl["d_out"], non_blocking=True)
                sl["ev_down"].record(s_dn)
        for s in (s_up, s_cp, s_dn):
            s.synchronize()

    run(depth * 2)
    torch.cuda.synchronize()
    # the pipeline must deliver the operator's results: compare the last download with a direct call
    last = slots[(depth * 2 - 1) % depth]
    chk = sets[(depth * 2 - 1) % n_host]
    ref_out = msda.ms_deform_attn_forward(chk["value"], shapes, start, chk["loc"], chk["attn"])
    got = views(last["h_out"], out_lay)[0]
    assert torch.allclose(got, ref_out.cpu(), rtol=1e-5, atol=1e-7), "e2e pipeline returned wrong results"
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(K)
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / K
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # the same pipeline with the kernels left out: what the host<->device link alone allows (all ranks copy at once)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    run(K, compute=False)
    torch.cuda.synchronize()
    ms_copy = (time.perf_counter() - t0) * 1e3 / K
    if world > 1:
        t = torch.tensor([ms_copy], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_copy = float(t.item())
    return {"value": w.samples * world / (ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "ms_per_step": ms, "steps": K,
            "copy_only_ms_per_step": ms_copy,
            "link_gbs_each_way_all_ranks": round(world * max(h2d, d2h) / (ms_copy * 1e-3) / 1e9, 1),
            "limit": ("host<->device copies: the same pipeline WITHOUT the kernels takes "
                      f"{ms_copy:.3f} ms/step ({100 * ms_copy / ms:.0f} % of the e2e step); H2D and D2H of all ranks share the host's PCIe / memory path"),
            "path": "pinned host arena -> 1 H2D -> ms_deform_attn_forward + ms_deform_attn_backward (C ABI) -> 1 D2H of out, "
                    "grad_value, grad_loc, grad_attn; 3-slot pipeline on 3 streams; host wall clock"}


def north_star_forward(torch, msda, dev, tdt, elt, peak):
    """Forward at BASELINE.json's target shape -- 4-level COCO 800x1333 pyramid, 300 queries, 8 heads, 4 points -- for the
    per-GPU batches of configs[4] (N = 2, "C5DEC") and configs[3] (N = 32, "C4DEC"): us per launch, Gsamples/s, fraction of
    the measured HBM peak by SURVEY 8(d) bytes (`frac`, counts the whole value tensor although 300 queries touch part of it)
    and by touched bytes (`frac_touched`, rows where a tap lands: the bytes that must move).  Device-resident rotating sets."""
    from aloception_oss_b200.synthetic import WORKLOADS, device_inputs

    res = {}
    for name in ("C5DEC", "C4DEC"):
        w = WORKLOADS[name]
        fb = w.algorithmic_bytes(elt, False)
        n_sets = max(2, min(8, int(3 * L2_BYTES / fb) + 2))
        sets = [device_inputs(w, seed=500 + i, device=dev, dtype=tdt, loc_mode="unit") for i in range(n_sets)]
        for s_ in sets:
            s_["out"] = torch.empty((w.N, w.Lq, w.M * w.D), dtype=tdt, device=dev)
        fn = lambda s: msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], out=s["out"])
        for i in range(3):
            fn(sets[i % n_sets])
        torch.cuda.synchronize()
        n = 64
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(n):
                fn(sets[i % n_sets])
        g.replay()
        torch.cuda.synchronize()
        times = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / n * 1e3)
        us = statistics.median(times)
        tb = touched_bytes(torch, w, sets[0], elt)
        res[name] = {"workload": workload_string(w, "unit"), "us_per_launch": round(us, 2),
                     "gsamples_per_s": round(w.samples / us / 1e3, 2), "algorithmic_bytes": fb,
                     "frac": round(fb / (us * 1e-6) / 1e9 / peak, 4), "touched_bytes": tb["fwd"],
                     "frac_touched": round(tb["fwd"] / (us * 1e-6) / 1e9 / peak, 4),
                     "value_rows_touched": tb["value_rows_touched"], "value_rows_total": tb["value_rows_total"]}
        del g, sets
        torch.cuda.empty_cache()
    return res


def north_star_census(torch, msda, dist, dev, tdt, world, rank, global_batch=32):
    """BASELINE.json configs[3] as far as this path goes: the 12 operator calls of one DeformableDETR-R50 inference forward
    -- 6 encoder calls (Lq = S = 22 223) + 6 decoder calls (Lq = 300) on the 800x1333 pyramid -- for a GLOBAL batch of 32
    images split over the ranks (strong scaling, no collective on the data path).  Each call has its own input set (a
    layer's value tensor is new), one CUDA graph of the 12 forwards, median of 5 replays, max over ranks."""
    from aloception_oss_b200.synthetic import WORKLOADS, device_inputs

    n_local = global_batch // world + (1 if rank < global_batch % world else 0)
    if n_local == 0:
        return None
    calls = []
    for name, mode, n_sets in (("C4ENC", "raster", 3), ("C4DEC", "unit", 3)):
        w = WORKLOADS[name].with_batch(n_local)
        sets = [device_inputs(w, seed=900 + 10 * rank + i, device=dev, dtype=tdt, loc_mode=mode) for i in range(n_sets)]
        for s_ in sets:
            s_["out"] = torch.empty((w.N, w.Lq, w.M * w.D), dtype=tdt, device=dev)
        calls += [sets[i % n_sets] for i in range(6)]
    fn = lambda s: msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], out=s["out"])
    for s_ in calls[:2] + calls[6:8]:
        fn(s_)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for s_ in calls:
            fn(s_)
    g.replay()
    torch.cuda.synchronize()
    times = []
    for _ in range(5):
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    t = torch.tensor([statistics.median(times)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    del g, calls
    torch.cuda.empty_cache()
    samples = global_batch * (22223 + 300) * 8 * 16 * 6
    return {"workload": "configs[3]: 6 encoder (Lq = S = 22223) + 6 decoder (Lq = 300) forward calls, 800x1333 pyramid, "
                        "global batch 32 split over the ranks", "global_batch": global_batch, "images_per_gpu": n_local,
            "ms_per_forward_census": round(ms, 3), "gsamples_per_s": round(samples / (ms * 1e-3) / 1e9, 2),
            "scaling": "strong", "timing": "one CUDA graph of the 12 calls, median of 5 replays, max over ranks"}


def training_census(torch, msda, dist, dev, tdt, world, per_gpu_batch=2):
    """BASELINE.json configs[4] as far as this path goes: the 12 forward + 12 backward operator calls of one DeformableDETR-R50
    training step at 2 images per GPU (weak scaling: every rank does the same work), 800x1333 pyramid.  One CUDA graph of the 24
    calls (each layer its own inputs and result buffers), median of 5 replays, max over ranks."""
    from aloception_oss_b200.synthetic import WORKLOADS, device_inputs

    calls = []
    for name, mode, n_sets in (("C5ENC", "raster", 6), ("C5DEC", "unit", 6)):
        w = WORKLOADS[name].with_batch(per_gpu_batch)
        for i in range(n_sets):
            s_ = device_inputs(w, seed=700 + i, device=dev, dtype=tdt, loc_mode=mode)
            s_["out"] = torch.empty((w.N, w.Lq, w.M * w.D), dtype=tdt, device=dev)
            s_["grads"] = [torch.empty_like(s_["value"]), torch.empty_like(s_["loc"]), torch.empty_like(s_["attn"])]
            calls.append(s_)
    fwd = lambda s: msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], out=s["out"])
    bwd = lambda s: msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"], grads=s["grads"])

    def step():
        for s_ in calls:
            fwd(s_)
        for s_ in reversed(calls):
            bwd(s_)

    step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    g.replay()
    torch.cuda.synchronize()
    times = []
    for _ in range(5):
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    t = torch.tensor([statistics.median(times)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    del g, calls
    torch.cuda.empty_cache()
    samples = per_gpu_batch * world * (22223 + 300) * 8 * 16 * 6
    return {"workload": "configs[4]: 12 forward + 12 backward operator calls of one training step, 800x1333 pyramid, "
                        "2 images per GPU", "images_per_gpu": per_gpu_batch, "global_batch": per_gpu_batch * world,
            "ms_per_step_census": round(ms, 3), "gsamples_per_s_fwd_bwd": round(samples / (ms * 1e-3) / 1e9, 2),
            "scaling": "weak", "timing": "one CUDA graph of the 24 calls, median of 5 replays, max over ranks"}


def run_extra(torch, msda, _capi, dev, tdt, elt, peak):
    """Other COCO shapes (not the headline): fwd and bwd us/launch and roofline fraction, device-resident."""
    from aloception_oss_b200.synthetic import WORKLOADS, device_inputs

    res = {}
    for name, mode in (("C4DEC", "unit"), ("C5DEC", "unit"), ("ENC", "raster"), ("C5ENC", "raster"), ("C4ENC", "raster")):
        w = WORKLOADS[name]
        step_bytes = w.algorithmic_bytes(elt, False) + w.algorithmic_bytes(elt, True)
        n_sets = max(2, min(8, int(2 * L2_BYTES / step_bytes) + 2))
        sets = [device_inputs(w, seed=77 + i, device=dev, dtype=tdt, loc_mode=mode) for i in range(n_sets)]
        out = {}
        for what, fn in (("fwd", lambda s: msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"])),
                         ("bwd", lambda s: msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"]))):
            for i in range(3):
                fn(sets[i % n_sets])
            torch.cuda.synchronize()
            n = 40
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(n):
                    fn(sets[i % n_sets])
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            del g
            us = e0.elapsed_time(e1) / n * 1e3
            b = w.algorithmic_bytes(elt, what == "bwd")
            out[what] = {"us": round(us, 2), "gsamples_per_s": round(w.samples / us / 1e3, 3),
                         "hbm_frac": round(b / (us * 1e-6) / 1e9 / peak, 4)}
        res[f"{name}[{mode}]"] = out
        del sets
        torch.cuda.empty_cache()
    return res


if __name__ == "__main__":
    main()
