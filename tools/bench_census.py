#!/usr/bin/env python3
"""Operator census of ONE DeformableDETR-R50 inference forward at BASELINE.json configs[3] (B=32 synthetic 800x1333 images,
batch-sharded over the GPUs of one box): 6 encoder self-attention calls (Lq = S = 22 223) + 6 decoder cross-attention calls
(Lq = 300) of the operator per forward, forward only, fp32.  STRONG scaling: the global batch of 32
is split over the ranks (N = 32 / world per GPU), no collective on the data path.

    python tools/bench_census.py                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 tools/bench_census.py

``--train``: BASELINE.json configs[4] instead -- the 6 + 6 forward AND backward calls of one training step at the reference's
per-GPU batch of 2 (B = 16 on 8 GPUs, WEAK scaling: every rank has its own 2 images; the DDP gradient all-reduce of the model is
outside the operator).

Prints one JSON line on rank 0 (device time, max over ranks)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import aloception_oss_b200 as msda
from aloception_oss_b200.synthetic import WORKLOADS, device_inputs

GLOBAL_BATCH = 32


def main():
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    msda.load_ops()
    train = "--train" in sys.argv
    n_local = 2 if train else GLOBAL_BATCH // world
    enc = WORKLOADS["C4ENC"].with_batch(n_local)
    dec = WORKLOADS["C4DEC"].with_batch(n_local)
    # two input sets per shape: 6 layers alternate between them (each layer of the real model has its own activations)
    enc_sets = [device_inputs(enc, seed=100 * rank + i, device=dev, loc_mode="raster") for i in range(2)]
    dec_sets = [device_inputs(dec, seed=100 * rank + 10 + i, device=dev, loc_mode="unit") for i in range(2)]
    fwd = lambda s: msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"])

    bwd = lambda s: msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"])

    def forward_pass():
        for i in range(6):
            fwd(enc_sets[i % 2])
        for i in range(6):
            fwd(dec_sets[i % 2])
        if train:
            for i in range(6):
                bwd(dec_sets[i % 2])
            for i in range(6):
                bwd(enc_sets[i % 2])

    for _ in range(3):
        forward_pass()
    torch.cuda.synchronize()
    reps = 10
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            forward_pass()
    g.replay()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    samples = (n_local * world) * 6 * (enc.Lq + dec.Lq) * enc.M * enc.L * enc.P
    if rank == 0:
        what = "fwd+bwd" if train else "forward"
        print(json.dumps({
            "metric": f"MSDeformAttn {what} Gsamples/s, DeformableDETR-R50 operator census (6 encoder + 6 decoder calls), "
                      f"B={n_local * world} 800x1333",
            "value": samples / (ms * 1e-3) / 1e9, "unit": "Gsamples/s", "n_gpus": world, "ms_per_step" if train else "ms_per_forward": ms,
            "scaling": "weak" if train else "strong", "per_gpu_batch": n_local, "dtype": "f32", "data": "synthetic",
            "launch": f"CUDA graph of {reps} passes ({36 if train else 12} launches each)",
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
