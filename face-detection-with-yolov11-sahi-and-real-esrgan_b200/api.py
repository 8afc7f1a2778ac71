"""Batch entry point of the fused path: many same-sized images in one call (what a serving loop or the evaluator's
dataset loop would use instead of calling get_sliced_prediction image by image)."""
from __future__ import annotations

import gc

from typing import List, Sequence

import numpy as np
import torch

from . import ops
from .sahi_api.annotation import Category
from .sahi_api.prediction import ObjectPrediction, PredictionResult

_FACE = Category(id=0, name="face")


def get_sliced_prediction_batch(images: Sequence, detection_model, slice_height: int, slice_width: int,
                                overlap_height_ratio: float = 0.2, overlap_width_ratio: float = 0.2,
                                perform_standard_pred: bool = True, postprocess_type: str = "GREEDYNMM",
                                postprocess_match_metric: str = "IOS", postprocess_match_threshold: float = 0.5,
                                postprocess_class_agnostic: bool = False, as_objects: bool = True, pool=None):
    """images: HWC uint8 arrays / CPU tensors (pinned for async H2D) of ONE common size — or encoded JPEGs (`bytes`) of one
    common size, decoded on the device (f4).  Returns a list of PredictionResult (as_objects=True) or the raw
    engine.DetectionBatch."""
    eng = detection_model.engine()
    eng.truncate = True
    if isinstance(images[0], (bytes, bytearray)):
        # (f4) encoded JPEGs: decoded on the device by nvJPEG straight into the image pool (RGB, like PIL's convert("RGB")).
        # Opt-in by passing bytes: the pixels differ from PIL's libjpeg by a few LSB, so this is outside the bit-exact claims.
        w, h, _ = ops.jpeg_info(bytes(images[0]))
        if pool is None or (pool.n, pool.h, pool.w) != (len(images), h, w):
            pool = ops.ImagePool(len(images), h, w, eng.device)
        for i, im in enumerate(images):
            pool.upload_jpeg(i, bytes(im))
    else:
        h, w = images[0].shape[:2]
        if pool is None or (pool.n, pool.h, pool.w) != (len(images), h, w):
            pool = ops.ImagePool(len(images), h, w, eng.device)
        for i, im in enumerate(images):
            pool.upload(i, im, non_blocking=True)
    batch = eng.detect(pool, slice_height, slice_width, overlap_height_ratio, overlap_width_ratio,
                       perform_standard_pred, postprocess_type, postprocess_match_metric,
                       postprocess_match_threshold, postprocess_class_agnostic)
    if not as_objects:
        return batch
    out: List[PredictionResult] = []
    gc_was_on = gc.isenabled()
    gc.disable()  # thousands of small acyclic objects: the cyclic collector's generation-0 passes would dominate
    try:
        for i in range(len(images)):
            boxes, scores, kpts, has_k = batch.image(i)
            ib, sc, hk = boxes.astype(np.int64).tolist(), scores.tolist(), has_k.tolist()
            preds = [ObjectPrediction.from_merged_row(b[0], b[1], b[2], b[3], sc[j], _FACE, kpts[j] if hk[j] else None)
                     for j, b in enumerate(ib)]
            out.append(PredictionResult(object_prediction_list=preds, image=images[i], durations_in_seconds={},
                                        image_size=(w, h)))
    finally:
        if gc_was_on:
            gc.enable()
    return out


class _Slot:
    def __init__(self, n, h, w, device, rows_cap, stream):
        self.shape, self.device = (n, h, w), device
        self.own_pool = None   # upload target, allocated on first use (device-resident batches bring their own pool)
        self.pool = None       # the pool the batch in flight reads
        self.stream = stream
        self.event = torch.cuda.Event()
        self.uploaded = torch.cuda.Event()
        self.t0 = torch.cuda.Event(enable_timing=True)   # device-side start / end of this batch's compute (stats only)
        self.t1 = torch.cuda.Event(enable_timing=True)
        self.h_off = torch.empty((n + 1,), dtype=torch.int32).pin_memory()
        self.h_rows = torch.empty((rows_cap, ops.ROW), dtype=torch.float32).pin_memory()
        self.h_cmax = torch.empty((2,), dtype=torch.int32).pin_memory()  # max candidates per entry, max stage-1 rows per image
        self.dev = None
        self.images = None


def predict_stream(batches, detection_model, slice_height: int, slice_width: int, overlap_height_ratio: float = 0.2,
                   overlap_width_ratio: float = 0.2, perform_standard_pred: bool = True,
                   postprocess_type: str = "GREEDYNMM", postprocess_match_metric: str = "IOS",
                   postprocess_match_threshold: float = 0.5, postprocess_class_agnostic: bool = False, depth: int = 3,
                   rows_per_image_hint: int = 256, stats: dict | None = None, use_graphs: bool = True,
                   overlap_post: bool = True):
    """Pipelined batch prediction: yields one list[PredictionResult] per batch of `batches` (an iterable of equal-length
    lists of same-sized HWC uint8 images, ideally pinned CPU tensors — or of `ops.ImagePool`s already resident on the
    device), in order.

    `depth` batches are in flight: the H2D upload of batch i+1 (copy stream) and the D2H + result-object construction of
    batch i-1 (host) overlap with the device pipeline of batch i (one compute stream; each batch has its own image pool).
    With `overlap_post` the latency-bound post-processing kernels (stage-1 NMS, merge, attach: ~2 ms on 32-384 CTAs) run on
    a third stream, overlapping the full-image pass and the next batch's backbone.  depth = 3 by default: under a bandwidth-saturating compute stream the 151 MB upload of a C2 batch takes ~18 ms instead
    of 3 ms, and with only two batches in flight it was issued too late to finish before the compute stream needed it
    (4.5 ms idle per batch, measured with `stats`).
    `stats`, if given, accumulates host seconds: "enqueue" (uploads + kernel launches), "wait" (blocked on the device),
    "build" (result-object construction), "device": seconds between the first and last kernel of each batch on the
    compute stream, and "device_idle": seconds between one batch's last and the next batch's first kernel (CUDA events)."""
    import time as _time

    if stats is not None:
        for key in ("enqueue", "wait", "build", "device", "device_idle"):
            stats.setdefault(key, 0.0)
    prev_t1 = [None]
    eng = detection_model.engine()
    eng.truncate = True
    graphs_before = eng.use_graphs
    eng.use_graphs = bool(use_graphs)  # the backbone chunks are replayed as CUDA graphs (static per-stream input buffers)
    slots, pending, k = None, [], 0

    def finish(slot):
        t_a = _time.perf_counter()
        slot.event.synchronize()
        t_b = _time.perf_counter()
        if stats is not None:
            slot.t1.synchronize()
            stats["device"] += slot.t0.elapsed_time(slot.t1) * 1e-3
            if prev_t1[0] is not None:  # compute-stream gap between the previous batch's last and this batch's first kernel
                stats["device_idle"] += max(0.0, prev_t1[0].elapsed_time(slot.t0)) * 1e-3
            prev_t1[0] = slot.t1
        n = slot.pool.n
        off = slot.h_off.numpy().copy()
        total = int(off[-1])
        eng.check_det_overflow(int(slot.h_cmax[1]), slot.dev["tables"]["det_cap"])
        # compared with the capacity THIS batch was enqueued with (eng.cap may have grown since, through another batch's redo)
        if int(slot.h_cmax[0]) > slot.dev["cap"]:  # a slice overflowed the candidate capacity: redo this batch synchronously
            batch = eng.detect(slot.pool, slice_height, slice_width, overlap_height_ratio, overlap_width_ratio,
                               perform_standard_pred, postprocess_type, postprocess_match_metric, postprocess_match_threshold)
            off, total = batch.offsets, int(batch.offsets[-1])
            host = np.concatenate([batch.boxes, batch.scores[:, None], (batch.has_keypoints.astype(np.float32) - 1)[:, None].view(np.float32),
                                   batch.keypoints.reshape(-1, 15)], 1) if total else np.zeros((0, 21), np.float32)
            srcs = np.where(batch.has_keypoints, 0, -1)
        else:
            if total <= slot.h_rows.shape[0]:
                host = slot.h_rows[:total].numpy()
            else:  # more detections than the pinned window: fetch the rest
                host = slot.dev["rows"][:total].cpu().numpy()
            srcs = host[:, 5].copy().view(np.int32) if total else np.zeros((0,), np.int32)
        w, h = slot.pool.w, slot.pool.h
        out = []
        boxes_all = host[:, :4].astype(np.int64).tolist()
        scores_all = host[:, 4].tolist()
        kp_all = host[:, 6:21].reshape(-1, 5, 3).copy()
        # a batch creates ~30 k small acyclic objects: with the cyclic collector enabled its generation-0 passes are 40 % of
        # the construction time (measured), so it is paused for the loop
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            for i in range(n):
                a, b = int(off[i]), int(off[i + 1])
                preds = [ObjectPrediction.from_merged_row(bx[0], bx[1], bx[2], bx[3], scores_all[j], _FACE,
                                                          kp_all[j] if srcs[j] >= 0 else None)
                         for j, bx in zip(range(a, b), boxes_all[a:b])]
                out.append(PredictionResult(object_prediction_list=preds, image=slot.images[i], durations_in_seconds={}, image_size=(w, h)))
        finally:
            if gc_was_on:
                gc.enable()
        slot.dev, slot.images = None, None
        if stats is not None:
            stats["wait"] += t_b - t_a
            stats["build"] += _time.perf_counter() - t_b
        return out

    for images in batches:
        resident = isinstance(images, ops.ImagePool)
        n, h, w = (images.n, images.h, images.w) if resident else (len(images), *images[0].shape[:2])
        if slots is None:
            # ONE compute stream for every batch (kernels of two batches never interleave: the conv -> epilogue pairs keep
            # their L2 reuse, and one set of captured graphs / static buffers serves all slots) + ONE copy stream on which
            # the next batch's images are uploaded while the current batch computes.  The streams live on the engine: its
            # captured backbone graphs are keyed by stream.
            streams = eng.__dict__.setdefault("_pipeline_streams", [])
            while len(streams) < 3:  # compute, copy (uploads), post (NMS / merge / attach / D2H of the previous batch)
                streams.append(torch.cuda.Stream(device=eng.device))
            slots = [_Slot(n, h, w, eng.device, n * rows_per_image_hint, streams[0]) for j in range(depth)]
        slot = slots[k % depth]
        k += 1
        if slot.dev is not None:
            yield finish(pending.pop(0))
        if slot.shape != (n, h, w):
            raise ValueError("predict_stream needs batches of one common size")
        t_e = _time.perf_counter()
        if resident:  # an ImagePool that already lives on the device: no upload, no host image behind the results
            slot.pool, slot.images = images, [None] * n
        else:
            if slot.own_pool is None:
                slot.own_pool = ops.ImagePool(n, h, w, eng.device)
            slot.pool, slot.images = slot.own_pool, images
        with torch.cuda.stream(streams[1]):  # this slot's pool is free: finish() waited for the batch that last read it
            if not resident:
                for i, im in enumerate(images):
                    slot.pool.upload(i, im, non_blocking=True)
            slot.uploaded.record(streams[1])
        with torch.cuda.stream(slot.stream):
            slot.stream.wait_event(slot.uploaded)
            if stats is not None:  # fresh events per batch: the previous batch's pair is still needed for the idle gap
                slot.t0, slot.t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                slot.t0.record(slot.stream)
            dev = eng.detect(slot.pool, slice_height, slice_width, overlap_height_ratio, overlap_width_ratio,
                             perform_standard_pred, postprocess_type, postprocess_match_metric,
                             postprocess_match_threshold, postprocess_class_agnostic, to_host=False,
                             post_stream=streams[2] if overlap_post else None)
            if stats is not None:
                slot.t1.record(slot.stream)  # end of this batch's work on the compute stream
        with torch.cuda.stream(dev.get("stream", slot.stream)):  # results live on the post stream
            slot.h_off.copy_(dev["offsets"], non_blocking=True)
            slot.h_rows.copy_(dev["rows"][: slot.h_rows.shape[0]], non_blocking=True)
            cmax = dev["count_s"].max() if dev["count_f"] is None else torch.maximum(dev["count_s"].max(), dev["count_f"].max())
            slot.h_cmax.copy_(torch.stack((cmax, dev["dmax"])).to(torch.int32), non_blocking=True)
            slot.event.record(dev.get("stream", slot.stream))
        slot.dev = dev
        pending.append(slot)
        if stats is not None:
            stats["enqueue"] += _time.perf_counter() - t_e
    while pending:
        yield finish(pending.pop(0))
    eng.use_graphs = graphs_before
