"""ncu driver: ONE launch of sort_vertices at 64 x 16384 polygons (256 MB of algorithmic bytes, cold cache).

    ncu --set full --clock-control none --import-source on -k regex:sortv -o gpurun_out/sortv python tools/prof_sortv.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aloception_oss_b200 import rotated_iou  # noqa: E402
from tools.bench_sortv import make  # noqa: E402

v, m, nv = make(64, 16384, 100, torch.device("cuda:0"))
torch.cuda.synchronize()
idx = rotated_iou.sort_v(v, m, nv)
torch.cuda.synchronize()
print(idx[0, 0].tolist())
