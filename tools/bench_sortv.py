"""Measure the sm_100a `sort_vertices` kernel on one B200: JSON lines on stdout.

    python tools/bench_sortv.py [--ref] [--cpu]

unit = one polygon (m = 24 candidates); algorithmic bytes per polygon = 24*8 (vertices) + 24 (mask) + 4 (num_valid) + 36
(indices) = 256 B, each tensor counted once.  Inputs resident in HBM; `sets` distinct input sets are rotated so that consecutive
launches never find their data in the 126 MB L2 (stated per line); CUDA events on the launch stream; the launches of one timed
region are captured in a CUDA graph.  --ref additionally times the reference's own kernel (oracle/_ref/sort_vertices_ref.so,
legacy default stream, eager launches: its figure includes launch overhead only where the kernel is shorter than a launch);
--cpu times the C oracle on a bounded sample (1 thread).  Prints roofline against MEASURED_PEAKS.json hbm_gbs."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from aloception_oss_b200 import rotated_iou  # noqa: E402
from bench import hbm_peak  # noqa: E402

BYTES_PER_POLYGON = 24 * 8 + 24 + 4 + 36


def make(b, n, seed, dev):
    g = torch.Generator().manual_seed(seed)
    v = torch.rand(b, n, 24, 2, generator=g)
    m = torch.rand(b, n, 24, generator=g) < 0.2
    # at most 8 valid (what two rectangles can produce, and what the reference kernel tolerates)
    extra = m.int().cumsum(-1) > 8
    m = m & ~extra
    nv = m.sum(-1).int()
    mean = (v * m[..., None]).sum(2, keepdim=True) / nv.clamp(min=1)[..., None, None]
    return (v - mean).to(dev).contiguous(), m.to(dev).contiguous(), nv.to(dev).contiguous()


def time_graph(fn_list, iters, dev):
    """fn_list: one callable per input set; returns average microseconds per launch."""
    s = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(s):
        for f in fn_list:
            f()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for f in fn_list:
                f()
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.synchronize()
        e0.record(s)
        for _ in range(iters):
            g.replay()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (iters * len(fn_list))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--variant", type=int, default=0, help="sortv_set_variant: 0 balanced tile kernel, 1 register kernel, 2 generic, 3 unbalanced tile kernel")
    args = ap.parse_args()
    assert torch.cuda.is_available(), "needs a CUDA device (no CPU fallback)"
    dev = torch.device("cuda:0")
    peak, peak_src = hbm_peak()
    L = rotated_iou.lib()
    rotated_iou.set_variant(args.variant)
    for name, b, n, sets in (("demo_8x1024", 8, 1024, 64), ("64x16384", 64, 16384, 4), ("1x1048576", 1, 1 << 20, 4)):
        data = [make(b, n, 100 + i, dev) for i in range(min(sets, 8))]
        while len(data) < sets:  # distinct buffers, repeated content
            data.append(tuple(t.clone() for t in data[len(data) % 8]))
        outs = [torch.empty(b, n, 9, dtype=torch.int32, device=dev) for _ in range(sets)]
        total = b * n

        def mk(i):
            v, m, nv = data[i]
            o = outs[i]
            return lambda: L.sortv_sort_vertices(v.data_ptr(), m.data_ptr(), nv.data_ptr(), o.data_ptr(), b, n, 24,
                                                 torch.cuda.current_stream().cuda_stream)

        us = time_graph([mk(i) for i in range(sets)], 20, dev)
        gbs = total * BYTES_PER_POLYGON / us * 1e-3
        line = dict(op="sort_vertices", variant=args.variant, workload=name, polygons=total, us=round(us, 2), gpolygons_per_s=round(total / us * 1e-3, 3),
                    roofline=dict(bound="hbm", achieved=round(gbs, 1), peak=peak, unit="GB/s", frac=round(gbs / peak, 3), peak_source=peak_src),
                    input_sets=sets, working_set_mb=round(sets * total * BYTES_PER_POLYGON / 1e6, 1))
        if args.ref:
            from oracle import build_ref_sortv

            if build_ref_sortv.built():
                import ctypes

                lib = ctypes.CDLL(build_ref_sortv.OUT_SO)
                fn = getattr(lib, build_ref_sortv.MANGLED)
                fn.restype = None
                fn.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 4
                ref_out = torch.empty(b, n, 9, dtype=torch.int32, device=dev)
                torch.cuda.synchronize()
                reps = 5 if total > 100000 else 50
                for w in range(2):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record(torch.cuda.default_stream())
                    for i in range(reps):
                        v, m, nv = data[i % sets]
                        fn(b, n, 24, v.data_ptr(), m.data_ptr(), nv.data_ptr(), ref_out.data_ptr())
                    e1.record(torch.cuda.default_stream())
                    torch.cuda.synchronize()
                ref_us = e0.elapsed_time(e1) * 1e3 / reps
                same = bool((ref_out == outs[(reps - 1) % sets]).all())
                line["reference_cuda_us"] = round(ref_us, 2)
                line["speedup_vs_reference_cuda"] = round(ref_us / us, 2)
                line["identical_to_reference_cuda"] = same
        if args.cpu:
            from oracle import sortv_oracle

            v, m, nv = (t.cpu().numpy() for t in data[0])
            take = min(b, max(1, 200000 // n))
            t0 = time.perf_counter()
            got = sortv_oracle.sort_vertices(v[:take], m[:take], nv[:take])
            dt = time.perf_counter() - t0
            line["cpu_oracle"] = dict(polygons=take * n, us=round(dt * 1e6, 1), gpolygons_per_s=round(take * n / dt * 1e-9, 5), cores=1,
                                      identical=bool((got == outs[0][:take].cpu().numpy()).all()))
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
