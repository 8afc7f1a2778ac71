#!/usr/bin/env python
"""Per-kernel micro-benchmarks: achieved GB/s against the algorithmic bytes of SURVEY §8(d), CUDA-event timed,
inputs larger than L2 (126 MB) so every launch streams from HBM.  One JSON line per measurement.

    python benchmarks/kernels.py [k1|k2|k3|k4|k5|ref|all]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fsd_b200  # noqa: E402,F401
from fsd_b200 import _cabi, ops  # noqa: E402

dev = torch.device("cuda:0")
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0


def timeit(fn, iters=10, warmup=3, reps=5):
    """Median / best device time of ONE call of fn in ms.  `reps` calls are captured into a CUDA graph and replayed so the
    figure is the kernel's device time, not the Python/ctypes launch path (which dominates for 20-us kernels)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    graph = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(reps):
                fn()
    except Exception:
        graph = None
        torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if graph is not None:
            graph.replay()
        else:
            for _ in range(reps):
                fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    return float(np.median(ts)), float(min(ts))


def report(name, nbytes, fn, **extra):
    med, best = timeit(fn)
    print(json.dumps({"kernel": name, "ms_median": round(med, 4), "ms_best": round(best, 4), "algorithmic_MB": round(nbytes / 1e6, 2),
                      "GBps_median": round(nbytes / med / 1e6, 1), "frac_of_measured_peak": round(nbytes / med / 1e6 / PEAK, 3),
                      "peak_GBps": PEAK, **extra}), flush=True)


def bench_ref():
    """The two reference points of this box: torch's device copy (the definition of MEASURED_PEAKS.hbm_gbs: bf16, read + write
    bytes) and a write-only fill (what a >= 94 %-write kernel like K1 competes with)."""
    n = 1 << 30
    a = torch.empty(n, dtype=torch.bfloat16, device=dev)
    b = torch.empty(n, dtype=torch.bfloat16, device=dev)
    report("torch copy_ 1 Gi bf16 (read+write)", 4 * n, lambda: b.copy_(a))
    report("torch fill_ 2 GiB (write only)", 2 * n, lambda: b.fill_(7))


def bench_k1():
    cases = [("C2 slices 512^2 -> 1024^2 (2x up)", 768, 1024, 512, 0.2, 32, False),
             ("C2 full 768x1024 (copy)", 768, 1024, None, 0.2, 64, True),
             ("C1 slices 640^2 -> 1024^2 (1.6x up)", 1080, 1920, 640, 0.2, 24, False),
             ("C1 full 1080x1920 -> 576x1024 (down)", 1080, 1920, None, 0.2, 64, True),
             ("C3 full 2160x3840 -> 576x1024 (down 3.75x)", 2160, 3840, None, 0.2, 16, True)]
    for name, H, W, sl, ov, N, full in cases:
        pool = ops.ImagePool(N, H, W, dev)
        pool.buf.random_(0, 256)
        if full:
            ent = torch.tensor([[i, 0, 0] for i in range(N)], dtype=torch.int32, device=dev)
            sw, sh = W, H
        else:
            boxes = _cabi.slice_plan(H, W, sl, sl, ov, ov)
            ent = torch.tensor([[i, b[0], b[1]] for i in range(N) for b in boxes], dtype=torch.int32, device=dev)
            sw = sh = sl
        g = _cabi.letterbox_geometry(sh, sw, 1024, 32)
        for dt, cl in ((torch.float16, False), (torch.float16, True), (torch.float32, False)):
            out = torch.empty((ent.shape[0], 3, g["out_h"], g["out_w"]), dtype=dt, device=dev,
                              memory_format=torch.channels_last if cl else torch.contiguous_format)
            nbytes = out.numel() * out.element_size() + N * H * W * 3
            report(f"K1 {name} {str(dt)[6:]} {'NHWC' if cl else 'NCHW'}", nbytes,
                   lambda: ops.gather_letterbox(pool, ent, sw, sh, 1024, 32, True, dt, out=out), entries=int(ent.shape[0]))
        del pool, out


def bench_k2():
    from fsd_b200.backbones.yolo11_pose import YOLO11Pose

    # class-logit statistics: (-6, 2) is bench.py's calibrated random head (0.24 % of the anchors pass conf 0.5, but 24 % pass
    # conf 0.01 — far more than a trained detector yields); (-9, 2) passes 1.4 % at conf 0.01, a realistic evaluator load
    for B, H, W, conf, cls_mean in ((96, 1024, 1024, 0.5, -6.0), (96, 1024, 1024, 0.01, -9.0), (96, 1024, 1024, 0.01, -6.0)):
        for cl in (True, False):
            g = torch.Generator(device=dev).manual_seed(0)
            levels = []
            for s in (8, 16, 32):
                h, w = H // s, W // s
                ts = [torch.randn((B, c, h, w), generator=g, device=dev, dtype=torch.float16) * sc + m
                      for c, sc, m in ((64, 1.5, 1.0), (1, 2.0, cls_mean), (15, 1.0, 0.0))]
                if cl:
                    ts = [t.contiguous(memory_format=torch.channels_last) for t in ts]
                levels.append(tuple(ts))
            A = YOLO11Pose.anchors_for(H, W)
            cand = torch.empty((B, 8192, ops.ROW), dtype=torch.float32, device=dev)
            count = torch.empty((B,), dtype=torch.int32, device=dev)
            ops.pose_decode(levels, conf, cand=cand, count=count)
            n = int(count.sum())
            full = B * 80 * A * 2 + n * ops.ROW * 4
            gated = B * A * 2 + n * (79 * 2 + ops.ROW * 4)
            report(f"K2a decode B={B} {H}x{W} conf={conf} cls~N({cls_mean:g},2) {'NHWC' if cl else 'NCHW'}", full,
                   lambda: ops.pose_decode(levels, conf, cand=cand, count=count), survivors=n, survivor_pct=round(100.0 * n / (B * A), 2),
                   gated_MB=round(gated / 1e6, 2))


def bench_k3():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_k3_merge_gpu import sahi_like_boxes

    rng = np.random.default_rng(0)
    sizes = ((256, 148), (1024, 148), (4096, 32), (9900, 8))
    if os.environ.get("FSD_K3_BENCH_SIZES"):  # e.g. "2048x4,4096x4,4096x32"
        sizes = tuple(tuple(int(v) for v in t.split("x")) for t in os.environ["FSD_K3_BENCH_SIZES"].split(","))
    for n_target, S in sizes:
        seg = sahi_like_boxes(rng, max(2, n_target // 3), dup=(1, 5), size=(10, 60), canvas=(3840, 2160))
        while len(seg) < n_target:
            seg = np.concatenate([seg, sahi_like_boxes(rng, 50, dup=(1, 5), size=(10, 60), canvas=(3840, 2160))])
        seg = seg[:n_target]
        rows = torch.from_numpy(np.tile(seg, (S, 1))).to(dev)
        offs = torch.arange(S, dtype=torch.int32, device=dev) * n_target
        for mtype, metric, kw in (("NMS", "IOU", dict(precision="fp64")), ("GREEDYNMM", "IOS", dict(precision="fp64")),
                                  ("NMS", "IOU", dict(precision="fp32", cmp_strict=True, max_keep=300, thr=0.7))):
            thr = kw.pop("thr", 0.5)
            fn = lambda: ops.merge_segments(rows, offs, None, n_target, merge_type=mtype, metric=metric, thr=thr, want_parent=False, **kw)  # noqa: E731
            r = fn()
            med, best = timeit(fn, iters=5, warmup=2)
            print(json.dumps({"kernel": f"K3 {mtype}/{metric} {kw['precision']} N={n_target} x{S} segments", "ms_median": round(med, 4),
                              "us_per_launch": round(med * 1e3, 1), "us_per_segment_amortised": round(med * 1e3 / S, 2),
                              "keeps_per_segment": int(r["keep_count"][0]), "mask_MB_if_materialised": round(n_target * ((n_target + 63) // 64) * 8 * 2 / 1e6, 3),
                              "pair_MFLOP": round(n_target * n_target / 2 * 20 / 1e6, 1)}), flush=True)


def bench_k4():
    """C5 shapes.  frames = 1 is the reference's per-image call (a 1080p frame is a 20 / 75 MB launch: latency-bound);
    frames = 8 is the batch mode (one launch over 8 frames of a stream)."""
    for H, W, scale, tile in ((1080, 1920, 2, 400), (1080, 1920, 4, 400)):
        for frames in (1, 8):
            R = 3  # rotate buffers so consecutive replays do not hit L2
            imgs = [torch.randint(0, 256, (frames, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(R)]
            table, _ = ops.esrgan_tile_table(H, W, scale, tile, 10, 0)
            tiles, tab_dev = ops.esrgan_crop(imgs[0], table, scale)
            crop_bytes = frames * (H * W * 3 + sum(3 * int(r[2]) * int(r[3]) for r in table) * 2)
            k = [0]

            def crop():
                k[0] = (k[0] + 1) % R
                ops.esrgan_crop(imgs[k[0]], table, scale, tab_dev=tab_dev, tiles=tiles)

            report(f"K4 crop {W}x{H} x{scale} tile {tile} fp16 frames={frames}", crop_bytes, crop, tiles=len(table))
            del imgs, tiles
            outs = [torch.rand(ops.esrgan_out_buffer(table, scale, torch.float16, dev, n_images=frames).shape, device=dev).half() for _ in range(R)]
            dst = [torch.empty((frames, H * scale, W * scale, 3), dtype=torch.uint8, device=dev) for _ in range(R)]
            stitch_bytes = frames * H * W * scale * scale * 3 * (2 + 1)

            def stitch():
                k[0] = (k[0] + 1) % R
                ops.esrgan_stitch(outs[k[0]], table, tab_dev, scale, H, W, out=dst[k[0]])

            report(f"K4 stitch {W}x{H} x{scale} tile {tile} fp16 frames={frames}", stitch_bytes, stitch, tiles=len(table))
            del outs, dst


def bench_k5():
    """conv epilogue passes of the backbone at the shapes of the C2 step (96 network inputs of 1024^2 per chunk)."""
    cl = lambda *shape: torch.randn(shape, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)  # noqa: E731
    for name, n, c, hw in (("b0 out 16ch 512^2", 96, 16, 512), ("b2.cv2 out 64ch 256^2", 96, 64, 256), ("b4.cv2 out 128ch 128^2", 96, 128, 128)):
        x = cl(n, c, hw, hw)
        bias = torch.randn((c,), device=dev, dtype=torch.float16)
        nbytes = 2 * x.numel() * 2
        report(f"K5 bias+SiLU in place {name}", nbytes, lambda: ops.bias_act_(x, bias, "silu"))
        report(f"K5 bias+SiLU general (dense out) {name}", nbytes, lambda: ops.bias_act(x, bias, "silu"))
        buf = cl(n, 2 * c, hw, hw)
        res = cl(n, c, hw, hw)
        report(f"K5 bias+SiLU+residual -> concat slot {name}", nbytes + x.numel() * 2,
               lambda: ops.bias_act(x, bias, "silu", out=buf[:, c:], residual=res))
        del x, buf, res
    buf = cl(96, 512, 32, 32)
    report("K5 SPPF pool 96 x 128ch 32x32 (read slot 0, write slots 1-3)", buf.numel() * 2, lambda: ops.sppf_pool_(buf))
    x = cl(96, 3, 1024, 1024)
    w = (torch.randn((16, 3, 3, 3), device=dev) * 0.4).half()
    bias = torch.randn((16,), device=dev).half()
    out = cl(96, 16, 512, 512)
    report("K6 stem conv3x3 s2 + bias + SiLU 96 x 3ch 1024^2 -> 16ch 512^2", (x.numel() + out.numel()) * 2, lambda: ops.stem_conv(x, w, bias, out=out))
    conv = lambda: ops.bias_act_(torch.nn.functional.conv2d(x, w.contiguous(memory_format=torch.channels_last), None, 2, 1), bias, "silu")  # noqa: E731
    torch.backends.cudnn.benchmark = True
    report("   (same layer as cuDNN conv + fsd_bias_act_inplace)", (x.numel() + out.numel()) * 2, conv)
    del x, out
    for name, k, n, hw in (("b2.cv1 32->32 256^2", 32, 32, 256), ("b2.cv2 48->64 256^2", 48, 64, 256), ("b4.cv2 96->128 128^2", 96, 128, 128),
                           ("head.cv3 64->64 128^2", 64, 64, 128), ("b6.cv1 128->128 64^2", 128, 128, 64)):
        x = cl(96, k, hw, hw)
        w = (torch.randn((n, k, 1, 1), device=dev) / k ** 0.5).half().contiguous(memory_format=torch.channels_last)
        bias = torch.randn((n,), device=dev).half()
        out = cl(96, n, hw, hw)
        nbytes = (x.numel() + out.numel()) * 2
        report(f"K7 pointwise conv + bias + SiLU {name}", nbytes, lambda: ops.pointwise_conv(x, w, bias, "silu", out=out))
        report("   (same layer as cuDNN conv + fsd_bias_act)", nbytes, lambda: ops.bias_act(torch.nn.functional.conv2d(x, w), bias, "silu", out=out))
        del x, out
    a, b = cl(96, 128, 64, 64), cl(96, 64, 128, 128)
    report("K5 upsample2x+concat 96 x (128ch 64^2 , 64ch 128^2)", (a.numel() + b.numel() + 96 * 192 * 128 * 128) * 2,
           lambda: ops.upsample2x_concat(a, b))


def bench_k10():
    """1x1 convolutions of the backbone (96 network inputs of 1024^2): tcgen05 kernel vs the mma.sync kernel vs cuDNN + epilogue."""
    cl = lambda *shape: torch.randn(shape, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)  # noqa: E731
    torch.backends.cudnn.benchmark = True
    for name, k, n, hw in (("b2.cv1 32->32 256^2", 32, 32, 256), ("b2.cv2 48->64 256^2", 48, 64, 256), ("head.cv3 64->64 128^2", 64, 64, 128),
                           ("b4.cv2 96->128 128^2", 96, 128, 128), ("h16.cv1 256->64 128^2", 256, 64, 128), ("b6.cv1 128->128 64^2", 128, 128, 64),
                           ("b6.cv2 192->128 64^2", 192, 128, 64), ("h13.cv1 384->128 64^2", 384, 128, 64), ("b8.cv1 256->256 32^2 (library only)", 256, 256, 32),
                           ("psa 128->256 32^2", 128, 256, 32)):
        x = cl(96, k, hw, hw)
        w = (torch.randn((n, k, 1, 1), device=dev) / k ** 0.5).half().contiguous(memory_format=torch.channels_last)
        bias = torch.randn((n,), device=dev).half()
        out = cl(96, n, hw, hw)
        nbytes = (x.numel() + out.numel()) * 2
        if ops.pointwise_tc_supported(k, n):
            os.environ.pop("FSD_K7_NO_TC", None)
            report(f"K10 tcgen05 pointwise conv + bias + SiLU {name}", nbytes, lambda: ops.pointwise_conv(x, w, bias, "silu", out=out))
            for key, val in (("FSD_SILU", "tanh"), ("FSD_K10_CTAS", "4"), ("FSD_K10_CTAS", "2"), ("FSD_K10_CTAS", "1")):
                os.environ[key] = val
                report(f"   K10 with {key}={val} {name}", nbytes, lambda: ops.pointwise_conv(x, w, bias, "silu", out=out))
                os.environ.pop(key)
            report(f"   K10 without activation {name}", nbytes, lambda: ops.pointwise_conv(x, w, bias, "none", out=out))
        os.environ["FSD_K7_NO_TC"] = "1"
        if ops.pointwise_conv_supported(k, n):
            report(f"   K7 mma.sync kernel {name}", nbytes, lambda: ops.pointwise_conv(x, w, bias, "silu", out=out))
        os.environ.pop("FSD_K7_NO_TC", None)
        report(f"   cuDNN conv + fsd_bias_act {name}", nbytes, lambda: ops.bias_act(torch.nn.functional.conv2d(x, w), bias, "silu", out=out))
        del x, out


def bench_conv3():
    """dense 3x3 stride-1 convolutions of the backbone (96 network inputs of 1024^2): fsd_conv3x3 (tcgen05 implicit GEMM) vs cuDNN + epilogue."""
    cl = lambda *shape: torch.randn(shape, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)  # noqa: E731
    torch.backends.cudnn.benchmark = True
    for name, k, n, hw in (("b2.m 16->8 256^2", 16, 8, 256), ("b4.m 32->16 128^2", 32, 16, 128), ("b4.m 16->32 128^2", 16, 32, 128),
                           ("head.cv2 64->64 128^2", 64, 64, 128), ("head.cv4 64->16 128^2", 64, 16, 128), ("c3k 32->32 64^2", 32, 32, 64),
                           ("b6 32->64 64^2", 32, 64, 64), ("b6 64->32 64^2", 64, 32, 64), ("head 64->64 64^2", 64, 64, 64),
                           ("head.cv4 128->16 64^2", 128, 16, 64), ("c3k 64->64 32^2", 64, 64, 32), ("head.cv4 256->16 32^2", 256, 16, 32),
                           ("head.cv4 16->16 128^2", 16, 16, 128), ("head.cv4 16->16 64^2", 16, 16, 64)):
        x = cl(96, k, hw, hw)
        w = (torch.randn((n, k, 3, 3), device=dev) / (3 * k ** 0.5)).half().contiguous(memory_format=torch.channels_last)
        taps = ops.conv3x3_tap_major(w)
        bias = torch.randn((n,), device=dev).half()
        out = cl(96, n, hw, hw)
        nbytes = (x.numel() + out.numel()) * 2
        report(f"K10 tcgen05 conv3x3 + bias + SiLU {name}", nbytes, lambda: ops.conv3x3(x, taps, bias, "silu", out=out))
        os.environ["FSD_K10_EW"] = "8"
        report(f"   with 8 epilogue warps {name}", nbytes, lambda: ops.conv3x3(x, taps, bias, "silu", out=out))
        os.environ.pop("FSD_K10_EW")
        report(f"   cuDNN conv + fsd_bias_act {name}", nbytes, lambda: ops.bias_act(torch.nn.functional.conv2d(x, w, None, 1, 1), bias, "silu", out=out))
        del x, out


def bench_dw():
    """depth-wise 3x3 layers of the head (96 network inputs of 1024^2): fsd_dwconv3x3 vs cuDNN grouped convolution + epilogue."""
    cl = lambda *shape: torch.randn(shape, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)  # noqa: E731
    torch.backends.cudnn.benchmark = True
    for name, c, hw in (("head.cv3 dw 64ch 128^2", 64, 128), ("head.cv3 dw 128ch 64^2", 128, 64), ("head.cv3 dw 64ch 64^2", 64, 64),
                        ("head.cv3 dw 256ch 32^2", 256, 32), ("head.cv3 dw 64ch 32^2", 64, 32)):
        x = cl(96, c, hw, hw)
        w = (torch.randn((c, 1, 3, 3), device=dev) / 3).half()
        taps = ops.dwconv3x3_tap_major(w)
        bias = torch.randn((c,), device=dev).half()
        out = cl(96, c, hw, hw)
        nbytes = (x.numel() + out.numel()) * 2
        report(f"K11 dwconv3x3 + bias + SiLU {name}", nbytes, lambda: ops.dwconv3x3(x, taps, bias, "silu", out=out))
        report(f"   cuDNN grouped conv + fsd_bias_act {name}", nbytes,
               lambda: ops.bias_act(torch.nn.functional.conv2d(x, w, None, 1, 1, 1, c), bias, "silu", out=out))
        del x, out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    for name, fn in (("ref", bench_ref), ("k1", bench_k1), ("k2", bench_k2), ("k3", bench_k3), ("k4", bench_k4), ("k5", bench_k5), ("k10", bench_k10), ("conv3", bench_conv3), ("dw", bench_dw)):
        if which in ("all", name):
            fn()
