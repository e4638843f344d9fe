#!/usr/bin/env python3
"""Diagnostic (GPU box, under torchrun): eager launch cost of one operator step before / after NCCL init and with the
NVML clock sampler running.  Prints one JSON line per rank to stderr."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import aloception_oss_b200 as msda
from aloception_oss_b200.synthetic import WORKLOADS, device_inputs

rank, local, world = (int(os.environ.get(k, "0")) for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
msda.load_ops()
w = WORKLOADS["C2"]
sets = [device_inputs(w, seed=i, device=dev) for i in range(4)]


def step(i):
    s = sets[i % 4]
    msda.ms_deform_attn_forward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"])
    msda.ms_deform_attn_backward(s["value"], s["shapes"], s["start"], s["loc"], s["attn"], s["grad_out"])


def measure(n=300):
    for i in range(20):
        step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        step(i)
    host = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) / n * 1e6
    return round(host, 1), round(total, 1)


res = {"rank": rank, "affinity": len(os.sched_getaffinity(0)), "cpu_count": os.cpu_count(),
       "omp": os.environ.get("OMP_NUM_THREADS"), "torch_threads": torch.get_num_threads()}
try:
    res["cpu.max"] = open("/sys/fs/cgroup/cpu.max").read().strip()
except Exception as e:
    res["cpu.max"] = repr(e)
res["before_init"] = measure()
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()
    res["after_init"] = measure()
    dist.barrier()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ClockSampler

c = ClockSampler(local).start()
res["with_sampler"] = measure()
c.stop()
res["after_sampler"] = measure()
sys.stderr.write(json.dumps(res) + "\n")
if world > 1:
    dist.destroy_process_group()
