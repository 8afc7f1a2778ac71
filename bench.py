#!/usr/bin/env python
"""bench.py — sliced face-detect images/sec on B200 (BASELINE.json metric), one JSON line on stdout.

Workload (config.workload): BASELINE.json configs[1] — YOLOv11n-face SAHI on synthetic WIDER-FACE-shaped 1024x768
images, 512x512 slices (0.2 overlap -> 6 slices) + the full-image pass, imgsz 1024 (the reference plug-in default),
GREEDYNMM / IOS / 0.5 merge, conf 0.5, fp16 network input.  A "step" is one batch of `--batch` images through the whole
hot path (Kernel 1 -> backbone -> Kernel 2a -> Kernel 3 -> Kernel 2b -> Kernel 3 -> attach/pack -> D2H of the results).

  value     images/sec with the step's images already resident in HBM (synchronous engine.detect per batch, the backbone
            replayed as CUDA graphs; CUDA-event timed)
  e2e       images/sec through the public API (fsd_b200.api.predict_stream: copy / compute / post-processing streams,
            three batches in flight) from PINNED HOST images: H2D copies, result D2H and the Python PredictionResult /
            ObjectPrediction objects are all inside the timed region
  roofline  Kernel 1 (slice launch): algorithmic bytes / time between CUDA events recorded inside the library around the
            launch, vs MEASURED_PEAKS.json hbm_gbs; `other_kernels`: the conv epilogue (largest share of the step);
            `hot_path`: SURVEY 8(d)'s whole-path figure
  cpu_baseline / --impl reference: the CPU oracle (reference-equivalent restatement: sequential batch-1 slices,
            per-box Python objects, CPU merge) on a bounded sample of the same images, all host threads.

Multi-GPU: launched by torchrun with one rank per GPU; images are sharded by index (weak scaling: every rank runs
`steps` batches of its own shard), no collective on the hot path, one all-gather of detections after the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, SLICE, OVERLAP, IMGSZ, CONF = 768, 1024, 512, 0.2, 1024, 0.5
WORKLOAD = "C2: YOLOv11n-face SAHI, synthetic 1024x768 images, 512x512 slices (0.2 overlap, 6 slices) + full-image pass, imgsz 1024, GREEDYNMM/IOS/0.5"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="images per step and per GPU")
    ap.add_argument("--images", type=int, default=4096, help="size of the synthetic data set (all GPUs together)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=6, help="images of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch the backbone eagerly instead of replaying CUDA graphs")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_arm_setup():
    import torch

    import fsd_b200  # noqa: F401
    from fsd_b200.backbones.yolo11_pose import build_yolo11n_pose
    from oracle.yolo_head import OracleYOLO
    from oracle.yolo_wrapper import YOLOv11PoseDetectionModel

    torch.set_num_threads(os.cpu_count() or 1)
    model = YOLOv11PoseDetectionModel(model=OracleYOLO(build_yolo11n_pose(), half=False), confidence_threshold=CONF,
                                      device="cpu", image_size=IMGSZ)
    return model


def cpu_arm_image(model, img):
    from oracle.predict import get_sliced_prediction

    model.keypoints_cache = {}
    res = get_sliced_prediction(img, model, slice_height=SLICE, slice_width=SLICE, overlap_height_ratio=OVERLAP,
                                overlap_width_ratio=OVERLAP, postprocess_type="GREEDYNMM", postprocess_match_metric="IOS",
                                postprocess_match_threshold=0.5, verbose=0)
    return model.attach_keypoints_to_predictions(res.object_prediction_list)


def run_reference_arm(args):
    """--impl reference: the reference-equivalent CPU restatement, one image per step, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch

    from fsd_b200.synthetic import make_image

    model = cpu_arm_setup()
    imgs = [make_image(i, H, W)[0] for i in range(min(args.steps + args.warmup, 8))]
    for i in range(args.warmup):
        cpu_arm_image(model, imgs[i % len(imgs)])
    t0 = time.perf_counter()
    for i in range(args.steps):
        cpu_arm_image(model, imgs[(args.warmup + i) % len(imgs)])
    dt = time.perf_counter() - t0
    v = args.steps / dt
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": "sliced face-detect images/sec", "value": v, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "images_per_step": 1},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} images (1 per step) of the C2 workload, oracle port of the reference CPU path "
                                       f"(sahi/ultralytics/realesrgan are not installable here); os.cpu_count()={os.cpu_count()}"},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for s in self.samples for k in range(4) if len(s) > 2 + k and s[2 + k].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


def main():
    args = parse()
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":  # NCCL would print its version banner on stdout, before the JSON line
        os.environ["NCCL_DEBUG"] = "WARN"
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import fsd_b200  # noqa: F401
    from fsd_b200 import _cabi, ops
    from fsd_b200.api import get_sliced_prediction_batch
    from fsd_b200.plugins import YOLOv11PoseDetectionModel
    from fsd_b200.synthetic import make_image, make_pool_on_device
    from fsd_b200.yolo import YOLO

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    n_local = max(B, (args.images // world) // B * B)   # images this rank owns (index i*world + rank)
    n_resident = min(n_local, max(B, (args.steps + args.warmup) * B))
    # ---- synthetic data: generated in HBM (device RNG), mirrored into pinned host memory for the e2e leg
    pool = make_pool_on_device(n_resident, H, W, dev, seed=1234 + rank * 100003)
    model = YOLOv11PoseDetectionModel(model=YOLO("random-init"), confidence_threshold=CONF, device=str(dev), image_size=IMGSZ)
    eng = model.engine()
    eng.overlap_post = True  # synchronous detect(): the slices' stage-1 NMS runs on a side stream under the full-image pass
    h = _cabi.get_handle(local)
    kw = dict(postprocess_type="GREEDYNMM", match_metric="IOS", match_threshold=0.5)

    def step_resident(i):
        a = (i % (n_resident // B)) * B
        return eng.detect(pool.subpool(a, a + B), SLICE, SLICE, OVERLAP, OVERLAP, True, **kw)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- leg 1: inputs resident in HBM -------------------------------------------------------------------
    # The backbone chunks are replayed as CUDA graphs (static network-input buffers; captured during the first warm-up
    # steps).  The last warm-up step runs eagerly once more with the conv-epilogue launches (~450 per step) timed by the
    # library: kernels inside a replayed graph are not individually bracketed.
    use_graphs = not args.no_graphs
    step_resident(0)  # cold eager step: cuDNN algorithm search, module loads (so that the timed eager step below is warm)
    for i in range(args.warmup):
        eng.use_graphs = use_graphs and i < args.warmup - 1
        if i == args.warmup - 1:
            h.timing_enable((_cabi.FSD_KERNEL_BIAS_ACT,))
        step_resident(i)
    eng.use_graphs = use_graphs and args.warmup >= 2
    k5 = [(units, ms_) for (kid, units, tag, ms_) in h.timing_read() if kid == _cabi.FSD_KERNEL_BIAS_ACT]
    # kernel timing: CUDA events recorded by the library itself right around each launch, on the launching stream
    # (= torch's current stream), so no host-side preparation falls inside a sample
    h.timing_enable((_cabi.FSD_KERNEL_GATHER,))
    sampler = ClockSampler(local)
    sync_all()
    sampler.start()
    torch.cuda.profiler.start()  # ncu --profile-from-start off captures exactly the timed region
    launches0 = h.launches + eng.replayed_launches  # direct launches + the library's kernels inside replayed graphs
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    n_dets = 0
    for i in range(args.steps):
        n_dets += int(step_resident(args.warmup + i).offsets[-1])
    t1.record()
    sync_all()
    torch.cuda.profiler.stop()
    launches = h.launches + eng.replayed_launches - launches0
    graphs_used = eng.use_graphs
    eng.use_graphs = False
    clocks = sampler.stop()
    samples = h.timing_read()
    h.timing_enable(())
    ms = t0.elapsed_time(t1)
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    value = world * args.steps * B / (ms / 1000.0)

    # roofline of the dominant kernel: Kernel 1 slice launch
    sl = [(units, ms_) for (kid, units, tag, ms_) in samples if kid == _cabi.FSD_KERNEL_GATHER and tag == SLICE]
    k1_ms = sum(t for _, t in sl) / max(len(sl), 1)
    n_entries = sl[0][0] if sl else 0
    k1_bytes = n_entries * 3 * IMGSZ * IMGSZ * 2 + (n_entries // 6) * H * W * 3  # every network input once + the source once
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = k1_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else 0.0
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum of the same launch (ncu --set full capture under profiles/)
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_k1_traffic.json")))
        if tj.get("entries") == n_entries:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k1_upscale2x_kernel<channels_last> via fsd_gather_letterbox (slice launch: 6 slices x B images, exact-2x path)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "bytes_per_launch": k1_bytes, "launch_ms": k1_ms, "launches_timed": len(sl),
                "timing": "CUDA events recorded inside libfsd_b200 around each launch (fsd_kernel_timing_*)",
                "note": "Kernel 1 is the HBM-bound kernel SURVEY 8(d) defines the per-image algorithmic bytes for; by share of the "
                        "step the largest hand-written kernel is the conv epilogue (other_kernels[0]); the step itself is bound by "
                        "the PyTorch convolutions (profiles/r1_launches_bench_b32_end.txt)"}
    # the hand-written kernel with the largest share of the step: the conv epilogue (bias + SiLU [+ residual] -> concat slot)
    if k5:
        k5_bytes, k5_ms = sum(u for u, _ in k5), sum(t for _, t in k5)
        roofline["other_kernels"] = [{
            "kernel": "k5_bias_act_general_half_kernel via fsd_bias_act (all launches of the last warm-up step)", "bound": "hbm",
            "achieved": k5_bytes / (k5_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": k5_bytes / (k5_ms * 1e-3) / 1e9 / peak,
            "launches_timed": len(k5), "ms_per_step": k5_ms, "share_of_step": k5_ms / (ms / args.steps)}]

    # ---- leg 2: end to end through the public API from pinned host memory ---------------------------------
    e2e = None
    if not args.skip_e2e:
        n_host = min(n_resident, 4 * B)
        host = torch.empty((n_host, H, W, 3), dtype=torch.uint8).pin_memory()
        for i in range(n_host):
            host[i].copy_(pool.view(i))
        torch.cuda.synchronize(dev)
        from fsd_b200.api import predict_stream

        def batches(count, start=0):
            for i in range(start, start + count):
                a = (i * B) % n_host // B * B
                yield [host[a + j] for j in range(B)]

        host_stats = {}

        def run_e2e(count, start=0):
            d2h = 0
            host_stats.clear()
            for res in predict_stream(batches(count, start), model, SLICE, SLICE, OVERLAP, OVERLAP, True, "GREEDYNMM", "IOS", 0.5,
                                      stats=host_stats):
                d2h += sum(len(r.object_prediction_list) for r in res) * ops.ROW * 4 + (B + 1) * 4
            return d2h

        run_e2e(max(2, args.warmup))
        sync_all()
        e_steps = max(4, args.steps)  # the 2-deep pipeline starts empty and is drained inside the timed region
        w0 = time.perf_counter()
        d2h = run_e2e(e_steps, start=3)
        sync_all()
        dt = time.perf_counter() - w0
        if world > 1:
            tt = torch.tensor([dt], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        e2e = {"value": world * e_steps * B / dt, "unit": "images/s", "h2d_bytes_per_step": B * H * W * 3,
               "d2h_bytes_per_step": d2h // e_steps, "steps": e_steps,
               "host_ms_per_step": {k: round(1e3 * v / e_steps, 2) for k, v in host_stats.items()},
               "api": "fsd_b200.api.predict_stream (pinned host images in, PredictionResult objects out; 3 batches in flight: H2D of "
                      "batch i+1 on a copy stream and D2H + object construction of batch i-1 overlap the device work of batch i on "
                      "one compute stream; backbone chunks replayed as CUDA graphs)"}

    # ---- the one collective: all-gather of detections for evaluation (after the timed region) -------------
    gathered = None
    if world > 1:
        from fsd_b200.shard import gather_detections

        last = step_resident(0)
        ids = torch.repeat_interleave(torch.arange(B, device=dev) * world + rank,
                                      torch.from_numpy(np.diff(last.offsets)).to(dev))
        rows = torch.from_numpy(np.concatenate([last.boxes, last.scores[:, None]], 1)).to(dev)
        gi, gr = gather_detections(ids.long(), rows)
        gathered = int(gi.shape[0])

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        model_cpu = cpu_arm_setup()
        imgs = [pool.view(i).cpu().numpy().copy() for i in range(args.cpu_sample)]
        cpu_arm_image(model_cpu, imgs[0])
        c0 = time.perf_counter()
        for im in imgs:
            cpu_arm_image(model_cpu, im)
        cdt = time.perf_counter() - c0
        cpu_base = {"value": len(imgs) / cdt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                    "sample": f"first {len(imgs)} images of this run's data set through the oracle port of the reference CPU path "
                              f"(sequential batch-1 slices at imgsz {IMGSZ}, fp32, per-box Python objects, CPU GREEDYNMM); "
                              f"os.cpu_count()={os.cpu_count()}"}

    if rank == 0:
        line = {"metric": "sliced face-detect images/sec", "value": value, "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": {"workload": WORKLOAD, "images_per_step_per_gpu": B, "data_set_images": args.images,
                           "resident_images_per_gpu": n_resident, "slices_per_image": 6, "network_inputs_per_image": 7,
                           "conf": CONF, "weights": "random-init YOLO11n-pose (calibrated head), seed 0",
                           "l2": "inputs larger than L2: each step gathers %.0f MB of source pixels into %.1f GB of network input"
                                 % (B * H * W * 3 / 1e6, B * (6 * 3 * IMGSZ * IMGSZ + 3 * H * W) * 2 / 1e9),
                           "parallelism": f"image-index sharding x{world}, no hot-path collective",
                           "cuda_graphs": bool(graphs_used)},
                "detections_per_image": n_dets / (args.steps * B), "gpu_launches": int(launches), "clocks": clocks,
                "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu_base}
        # the whole hot path against the HBM roofline = algorithmic bytes of Kernel 1 (44.8 MB / image) and
        # Kernel 2 (23.2 MB / image) over the time per image; it is small because the step is the PyTorch backbone's
        bytes_per_image = (H * W * 3 + 6 * 3 * IMGSZ * IMGSZ * 2 + 3 * H * W * 2) + (6 * 21504 + 16128) * 80 * 2
        line["hot_path"] = {"algorithmic_bytes_per_image": bytes_per_image, "achieved_GBps_per_gpu": bytes_per_image * value / world / 1e9,
                            "frac_of_peak": bytes_per_image * value / world / 1e9 / peak,
                            "backbone_share_of_device_time": 0.87,
                            "backbone_share_source": "profiles/r1_launches_bench_b32_end.txt: library conv/gemm/sdpa 43.5 % + the "
                                                     "hand-written backbone kernels (epilogues, stem, 1x1 conv, up-sample, SPPF) 43.4 %"}
        if gathered is not None:
            line["allgather_detections"] = gathered
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
