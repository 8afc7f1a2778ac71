#!/usr/bin/env python
"""Run ONE kernel configuration a few times (the target of `ncu --set full -k regex:...`).
    python benchmarks/one_kernel.py k1_c2 | k1_c2_f32 | k2 | k3 | k4"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fsd_b200  # noqa: E402,F401
from fsd_b200 import _cabi, ops  # noqa: E402

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "k1_c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
if which in ("k1_c2", "k1_c2_f32"):
    N = 32
    pool = ops.ImagePool(N, 768, 1024, dev)
    pool.buf.random_(0, 256)
    boxes = _cabi.slice_plan(768, 1024, 512, 512, 0.2, 0.2)
    ent = torch.tensor([[i, b[0], b[1]] for i in range(N) for b in boxes], dtype=torch.int32, device=dev)
    dt = torch.float32 if which.endswith("f32") else torch.float16
    out = torch.empty((ent.shape[0], 3, 1024, 1024), dtype=dt, device=dev)
    for _ in range(iters):
        ops.gather_letterbox(pool, ent, 512, 512, 1024, 32, True, dt, out=out)
elif which == "k1_nhwc":
    N = int(os.environ.get("FSD_N", "32"))
    pool = ops.ImagePool(N, 768, 1024, dev)
    pool.buf.random_(0, 256)
    boxes = _cabi.slice_plan(768, 1024, 512, 512, 0.2, 0.2)
    ent = torch.tensor([[i, b[0], b[1]] for i in range(N) for b in boxes], dtype=torch.int32, device=dev)
    out = torch.empty((ent.shape[0], 3, 1024, 1024), dtype=torch.float16, device=dev, memory_format=torch.channels_last)
    for _ in range(iters):
        ops.gather_letterbox(pool, ent, 512, 512, 1024, 32, True, torch.float16, out=out)
elif which == "k1_c1":  # C1 slices 640^2 -> 1024^2: the sixteenths path
    N = 24
    pool = ops.ImagePool(N, 1080, 1920, dev)
    pool.buf.random_(0, 256)
    boxes = _cabi.slice_plan(1080, 1920, 640, 640, 0.2, 0.2)
    ent = torch.tensor([[i, b[0], b[1]] for i in range(N) for b in boxes], dtype=torch.int32, device=dev)
    out = torch.empty((ent.shape[0], 3, 1024, 1024), dtype=torch.float16, device=dev, memory_format=torch.channels_last)
    for _ in range(iters):
        ops.gather_letterbox(pool, ent, 640, 640, 1024, 32, True, torch.float16, out=out)
elif which in ("k10", "k10_wide"):  # 1x1 convolution on tcgen05: 32->32 at 256^2 (epilogue-heavy) / 96->128 at 128^2
    k, n, hw = (32, 32, 256) if which == "k10" else (96, 128, 128)
    x = torch.randn((96, k, hw, hw), device=dev).half().contiguous(memory_format=torch.channels_last)
    w = (torch.randn((n, k, 1, 1), device=dev) / k ** 0.5).half()
    b = torch.randn((n,), device=dev).half()
    out = torch.empty((96, n, hw, hw), device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
    for _ in range(iters):
        ops.pointwise_conv(x, w, b, "silu", out=out)
elif which in ("conv3", "conv3_small"):  # 3x3 convolution on tcgen05 (halo mode): 64->64 at 128^2 / 16->8 at 256^2
    k, n, hw = (64, 64, 128) if which == "conv3" else (16, 8, 256)
    x = torch.randn((96, k, hw, hw), device=dev).half().contiguous(memory_format=torch.channels_last)
    w = (torch.randn((n, k, 3, 3), device=dev) / (3 * k ** 0.5)).half()
    taps = ops.conv3x3_tap_major(w)
    b = torch.randn((n,), device=dev).half()
    out = torch.empty((96, n, hw, hw), device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
    for _ in range(iters):
        ops.conv3x3(x, taps, b, "silu", out=out)
elif which == "conv2":  # layer 1 on the space-to-depth stem output: 2x2 convolution, 64 -> 32, 257 x 257 -> 256 x 256
    x = torch.randn((32, 64, 257, 257), device=dev).half().contiguous(memory_format=torch.channels_last)
    w = (torch.randn((32, 64, 2, 2), device=dev) / 16).half()
    taps = w.permute(2, 3, 0, 1).contiguous()
    b = torch.randn((32,), device=dev).half()
    for _ in range(iters):
        out = ops.conv2x2(x, taps, b, "silu")
    torch.cuda.synchronize()
    ref = torch.nn.functional.silu(torch.nn.functional.conv2d(x.float(), w.float(), b.float()))
    print("conv2 max abs diff", float((out.float() - ref).abs().max()))
elif which == "k7":
    x = torch.randn((96, 48, 256, 256), device=dev).half().contiguous(memory_format=torch.channels_last)
    w = (torch.randn((64, 48, 1, 1), device=dev) / 7).half()
    b = torch.randn((64,), device=dev).half()
    out = torch.empty((96, 64, 256, 256), device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
    for _ in range(iters):
        ops.pointwise_conv(x, w, b, "silu", out=out)
elif which == "k3_big":  # config 3: 9900 boxes per image -> the cluster kernel
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_k3_merge_gpu import sahi_like_boxes

    rng = np.random.default_rng(0)
    seg = sahi_like_boxes(rng, 5000, dup=(1, 4), size=(10, 40), canvas=(3840, 2160))[:9900]
    S = 8
    rows = torch.from_numpy(np.tile(seg, (S, 1))).to(dev)
    offs = torch.arange(S, dtype=torch.int32, device=dev) * len(seg)
    for _ in range(iters):
        ops.merge_segments(rows, offs, None, len(seg), merge_type="GREEDYNMM", metric="IOS", thr=0.5, want_parent=False)
elif which == "k6":
    x = torch.rand((96, 3, 1024, 1024), device=dev).half().contiguous(memory_format=torch.channels_last)
    w = (torch.randn((16, 3, 3, 3), device=dev) * 0.4).half()
    b = torch.randn((16,), device=dev).half()
    out = torch.empty((96, 16, 512, 512), device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
    for _ in range(iters):
        ops.stem_conv(x, w, b, out=out)
elif which == "k5":
    x = torch.randn((96, 64, 256, 256), device=dev).half().contiguous(memory_format=torch.channels_last)
    buf = torch.empty((96, 128, 256, 256), device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
    res = torch.randn((96, 64, 256, 256), device=dev).half().contiguous(memory_format=torch.channels_last)
    b = torch.randn((64,), device=dev).half()
    for _ in range(iters):
        ops.bias_act(x, b, "silu", out=buf[:, 64:], residual=res)
        ops.bias_act(x, b, "silu")
elif which == "k2":
    B, H, W = 96, 1024, 1024
    g = torch.Generator(device=dev).manual_seed(0)
    levels = []
    for s_ in (8, 16, 32):
        h, w = H // s_, W // s_
        levels.append(tuple((torch.randn((B, c, h, w), generator=g, device=dev, dtype=torch.float16) * sc + m).contiguous(memory_format=torch.channels_last)
                            for c, sc, m in ((64, 1.5, 1.0), (1, 2.0, -6.0), (15, 1.0, 0.0))))
    cand = torch.empty((B, 4096, ops.ROW), dtype=torch.float32, device=dev)
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    for _ in range(iters):
        ops.pose_decode(levels, 0.5, cand=cand, count=count)
elif which == "k3":
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_k3_merge_gpu import sahi_like_boxes

    rng = np.random.default_rng(0)
    seg = sahi_like_boxes(rng, 400, dup=(1, 5), size=(10, 60), canvas=(3840, 2160))[:1024]
    S = 148
    rows = torch.from_numpy(np.tile(seg, (S, 1))).to(dev)
    offs = torch.arange(S, dtype=torch.int32, device=dev) * len(seg)
    for _ in range(iters):
        ops.merge_segments(rows, offs, None, len(seg), merge_type="GREEDYNMM", metric="IOS", thr=0.5, want_parent=False)
elif which == "k4":
    H, W, scale, tile = 1080, 1920, 2, 400
    frames = int(os.environ.get("FSD_FRAMES", "1"))
    img = torch.randint(0, 256, (frames, H, W, 3), dtype=torch.uint8, device=dev)
    table, _ = ops.esrgan_tile_table(H, W, scale, tile, 10, 0)
    for _ in range(iters):
        tiles, tab_dev = ops.esrgan_crop(img, table, scale)
        outb = torch.rand(ops.esrgan_out_buffer(table, scale, torch.float16, dev, n_images=frames).shape, device=dev).half()
        ops.esrgan_stitch(outb, table, tab_dev, scale, H, W)
torch.cuda.synchronize()
print("done", which)
