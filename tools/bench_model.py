#!/usr/bin/env python3
"""Model-level benchmark of BASELINE.json configs[3] / configs[4]: the reference's OWN DeformableDETR-R50 (unmodified
classes from the bundle oracle/_ref/aloception_src, loaded by tools/ref_model.py), random-init weights, synthetic 800x1333
images, with every MSDeformAttn running on the B200 operator (aloception_oss_b200.integration.install()).

    python tools/bench_model.py --mode infer                  # configs[3]: B = 32 inference, batch-sharded over the ranks
    python tools/bench_model.py --mode train                  # configs[4]: B = 2 per GPU, fwd + loss + bwd (+ AdamW step)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
        tools/bench_model.py --mode infer

Multi-GPU (one process per GPU, NCCL): rank 0's random-init parameters and buffers are BROADCAST once (one flat buffer,
~160 MB: SURVEY.md section 8e) so that every rank holds the same model; inference shards the global batch (strong scaling,
no collective on the data path); training wraps the model in DistributedDataParallel (bucketed gradient all-reduce,
alonet/common/pl_helpers.py:372 strategy "ddp") and uses the reference's criterion with its scalar all_reduce(num_boxes)
(alonet/detr/criterion.py:411-413).  Time: CUDA events, barrier + synchronize on both sides, max over ranks.  The operator's
share is measured with event pairs around each of its 12 (+12 backward) calls per step.

Prints one JSON line on rank 0.  What is measured is the REFERENCE model; nothing of it is rebuilt in this repository.
"""
import argparse
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="infer", choices=["infer", "train"])
    ap.add_argument("--global-batch", type=int, default=32, help="infer: images over all ranks")
    ap.add_argument("--per-gpu-batch", type=int, default=2, help="train: images per rank (reference default, data2detr.py:105)")
    ap.add_argument("--height", type=int, default=800)
    ap.add_argument("--width", type=int, default=1333)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--boxes", type=int, default=7, help="train: target boxes per image")
    ap.add_argument("--no-optimizer", action="store_true")
    ap.add_argument("--autocast", default="off", choices=["off", "bf16", "f16"],
                    help="run the model under torch.autocast: the reference module then hands the operator 16-bit value next to fp32 "
                         "sampling locations / attention weights (MSDA_LOC_F32 | MSDA_ATTN_F32)")
    ap.add_argument("--unit-bwd", action="store_true", help="(knob) keep the unit-ordered backward for the encoder calls")
    ap.add_argument("--tile-bwd", action="store_true", help="(knob) tile-binned backward for the encoder calls")
    args = ap.parse_args()

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warnings.filterwarnings("ignore")
    from tools import ref_model

    alonet, aloscene = ref_model.load()
    from aloception_oss_b200 import _capi, functions

    if args.tile_bwd:
        _capi.set_tuning("bwd_tile_mode", 2)
    from alonet.deformable_detr import DeformableDetrR50

    # ---- the reference model, random init; ranks start DIFFERENT and are made equal by one broadcast from rank 0 ----
    torch.manual_seed(1234 + 1000 * rank)
    model = DeformableDetrR50(num_classes=91, device=dev)
    tensors = [p.data for p in model.parameters()] + [b.data for b in model.buffers() if b.is_floating_point()]
    n_params = sum(p.numel() for p in model.parameters())
    bcast = {"bytes": 0, "ms": 0.0}
    if world > 1:
        flat = torch._utils._flatten_dense_tensors(tensors)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        dist.broadcast(flat, src=0)
        torch.cuda.synchronize()
        bcast = {"bytes": flat.numel() * flat.element_size(), "ms": (time.perf_counter() - t0) * 1e3, "collective": "ncclBroadcast, one flat buffer"}
        for t, f in zip(tensors, torch._utils._unflatten_dense_tensors(flat, tensors)):
            t.copy_(f)
        chk = torch.stack([flat.double().sum(), flat.double().abs().sum()])
        ref = chk.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(chk, ref), "ranks hold different weights after the broadcast"
        del flat
    n_attn = sum(1 for m in model.modules() if type(m).__name__ == "MSDeformAttn")
    attn_cls = next(m for m in model.modules() if type(m).__name__ == "MSDeformAttn")
    assert type(attn_cls).__module__.startswith("alonet."), "the benchmark must run the REFERENCE module class"

    # ---- operator accounting: event pairs around every operator call ----
    log = []
    fwd0, bwd0 = functions.ms_deform_attn_forward, functions.ms_deform_attn_backward

    def timed(fn, kind):
        def wrap(value, *a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(value, *a, **k)
            e1.record()
            loc = a[2]
            op_dtypes.add("value %s, sampling_loc %s, attn_weight %s" % (str(value.dtype).replace("torch.", ""), str(loc.dtype).replace("torch.", ""),
                                                                         str(a[3].dtype).replace("torch.", "")))
            log.append((kind, e0, e1, loc.shape[0] * loc.shape[1] * loc.shape[2] * loc.shape[3] * loc.shape[4]))
            return r
        return wrap

    from aloception_oss_b200 import torch_ops

    if torch_ops.using_shim():  # the reference's MSDeformAttnFunction looks the ops up on this namespace at call time
        ns = torch.ops.alonet_custom
        ns.ms_deform_attn_forward = timed(ns.ms_deform_attn_forward, "fwd")
        ns.ms_deform_attn_backward = timed(ns.ms_deform_attn_backward, "bwd")
    else:
        functions.ms_deform_attn_forward = timed(fwd0, "fwd")
        functions.ms_deform_attn_backward = timed(bwd0, "bwd")

    # ---- synthetic input: aloscene.Frame batch with a padding mask (all valid), like the reference's data pipeline ----
    n_local = args.global_batch // world if args.mode == "infer" else args.per_gpu_batch
    assert n_local >= 1, "global batch smaller than the number of ranks"
    g = torch.Generator().manual_seed(77 + rank)
    names = [str(i) for i in range(91)]
    frames = []
    for i in range(n_local):
        f = aloscene.Frame(torch.rand(3, args.height, args.width, generator=g), names=("C", "H", "W")).norm_resnet()
        if args.mode == "train":
            nb = args.boxes
            cxcy = torch.rand(nb, 2, generator=g) * 0.6 + 0.2
            wh = torch.rand(nb, 2, generator=g) * 0.2 + 0.05
            labels = aloscene.Labels(torch.randint(0, 91, (nb,), generator=g).float(), encoding="id", labels_names=names, names=("N",))
            f.append_boxes2d(aloscene.BoundingBoxes2D(torch.cat([cxcy, wh], -1), boxes_format="xcyc", absolute=False,
                                                      labels=labels, names=("N", None)))
        frames.append(f)
    batch = aloscene.Frame.batch_list(frames).to(dev)

    import contextlib

    def cast():
        if args.autocast == "off":
            return contextlib.nullcontext()
        return torch.autocast("cuda", dtype=torch.bfloat16 if args.autocast == "bf16" else torch.float16)

    def to_f32(o):
        if torch.is_tensor(o):
            return o.float() if o.is_floating_point() else o
        if isinstance(o, dict):
            return {k: to_f32(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return type(o)(to_f32(v) for v in o)
        return o

    op_dtypes = set()
    if args.mode == "infer":
        model.eval()

        def step():
            with torch.no_grad(), cast():
                out = model(batch)
            return out["pred_logits"]
    else:
        from alonet.deformable_detr import DeformableCriterion, DeformableDetrHungarianMatcher

        model.train()
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
        crit = DeformableCriterion(matcher=DeformableDetrHungarianMatcher(1, 5, 2), loss_label_weight=1, loss_boxes_weight=5,
                                   loss_giou_weight=2, losses=["labels", "boxes"], aux_loss_stage=6, eos_coef=0.1)
        opt = None if args.no_optimizer else torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-4)

        def step():
            if opt is not None:
                opt.zero_grad(set_to_none=True)
            else:
                for p in model.parameters():
                    p.grad = None
            with cast():
                out = net(batch)
            if args.autocast != "off":  # the criterion (focal / L1 / GIoU losses, scipy matcher) runs in fp32, outside autocast
                out = to_f32(out)
            loss, _ = crit(out, batch)
            loss.backward()
            if opt is not None:
                opt.step()
            return loss.detach()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, args.warmup)):
        last = step()
    barrier()
    assert torch.isfinite(last).all(), "non-finite model output"
    log.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    wall = (time.perf_counter() - t0) * 1e3 / args.steps
    ms = e0.elapsed_time(e1) / args.steps
    op_ms = {"fwd": 0.0, "bwd": 0.0}
    op_samples = {"fwd": 0, "bwd": 0}
    calls = {"fwd": 0, "bwd": 0}
    for kind, a, b, ns in log:
        op_ms[kind] += a.elapsed_time(b)
        op_samples[kind] += ns
        calls[kind] += 1
    stats = torch.tensor([ms, wall, op_ms["fwd"] / args.steps, op_ms["bwd"] / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ms, wall, of, ob = (float(v) for v in stats.tolist())
    if rank == 0:
        images = n_local * world
        line = {
            "bench": "reference DeformableDetrR50 (unmodified alonet classes) on the B200 operator",
            "config": ("BASELINE.json configs[3]: inference, B=%d synthetic %dx%d, batch-sharded" % (images, args.height, args.width)) if args.mode == "infer"
            else ("BASELINE.json configs[4]: training step fwd + criterion + bwd%s, B=%d (%d per GPU), DDP" % ("" if args.no_optimizer else " + AdamW", images, n_local)),
            "mode": args.mode, "n_gpus": world, "global_batch": images, "per_gpu_batch": n_local,
            "dtype": "f32" if args.autocast == "off" else "torch.autocast(%s)" % args.autocast, "operator_dtypes": sorted(op_dtypes),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "ms_per_step_wall": round(wall, 3),
            "images_per_s": round(images / (ms * 1e-3), 2),
            "params_M": round(n_params / 1e6, 2), "msdeformattn_modules": n_attn, "module_class": type(attn_cls).__module__ + "." + type(attn_cls).__name__,
            "operator": {"calls_per_step": {k: v // args.steps for k, v in calls.items()}, "fwd_ms_per_step": round(of, 3), "bwd_ms_per_step": round(ob, 3),
                         "share_of_step": round((of + ob) / ms, 4),
                         "gsamples_per_s_inside_calls": round((op_samples["fwd"] + op_samples["bwd"]) / args.steps / max((of + ob) * 1e-3, 1e-9) / 1e9, 3),
                         "how": "CUDA event pairs around ms_deform_attn_forward / ms_deform_attn_backward (C ABI), max over ranks"},
            "weight_broadcast": bcast, "kernel_launches_libmsda": int(_capi.kernel_launch_count()),
            "bwd_schedule": "tile" if args.tile_bwd else "unit", "op_registration": "C++ shim" if torch_ops.using_shim() else "python",
            "scaling": "strong (global batch fixed)" if args.mode == "infer" else "weak (per-GPU batch fixed)",
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
