#!/usr/bin/env python
"""C5's enhancement stage on one GPU: 1920x1080 frames -> Real-ESRGAN x2 (RRDBNet, 23 blocks, random-init, fp16, tile 400,
pad 10) -> 3840x2160, through fsd_b200.enhancer.RealESRGANer.enhance_device (Kernel 4 crop / stitch around the PyTorch
network).  Prints frames/s and the share of the two Kernel 4 launches.  One JSON line."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fsd_b200  # noqa: E402,F401
from fsd_b200 import ops  # noqa: E402
from fsd_b200.backbones.rrdbnet import RRDBNet  # noqa: E402
from fsd_b200.enhancer import RealESRGANer  # noqa: E402

dev = torch.device("cuda:0")
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
torch.backends.cudnn.benchmark = True
net = RRDBNet(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=23, num_grow_ch=32, scale=2)
up = RealESRGANer(scale=2, model=net, tile=400, tile_pad=10, pre_pad=0, half=True, max_tile_batch=15)
x = torch.randint(0, 256, (frames, 1080, 1920, 3), dtype=torch.uint8, device=dev)
for _ in range(2):
    y = up.enhance_device(x)
torch.cuda.synchronize()
t = time.perf_counter()
n = 3
for _ in range(n):
    y = up.enhance_device(x)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / n
# Kernel 4 alone on the same shapes
table, _ = ops.esrgan_tile_table(1080, 1920, 2, 400, 10, 0)
tiles, tab = ops.esrgan_crop(x, table, 2)
outs = ops.esrgan_out_buffer(table, 2, torch.float16, dev, n_images=frames).normal_()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record(); ops.esrgan_crop(x, table, 2, tab_dev=tab, tiles=tiles); e1.record(); ops.esrgan_stitch(outs, table, tab, 2, 1080, 1920); e2.record()
torch.cuda.synchronize()
print(json.dumps({"workload": f"C5 enhancement stage: {frames} x 1920x1080 -> x2, RRDBNet-23 fp16, tile 400", "frames_per_s": frames / dt,
                  "ms_per_frame": 1e3 * dt / frames, "k4_crop_ms": e0.elapsed_time(e1), "k4_stitch_ms": e1.elapsed_time(e2),
                  "k4_share": (e0.elapsed_time(e2) * 1e-3) / dt, "out_shape": list(y.shape)}))
