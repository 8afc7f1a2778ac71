"""Batched device pipeline of the sliced-inference hot path:

    K1 gather+letterbox (slices, full image)  ->  PyTorch backbone (all slices of all images in one batch)
    ->  K2a decode+gate+compact  ->  K3 stage 1 (per-slice NMS, torchvision rule)  ->  K2b finalize (un-letterbox,
    int(), shift)  ->  K3 stage 2 (per-image NMS | GREEDYNMM | NMM)  ->  key-point attach  ->  pack  ->  ONE D2H.

The reference does the same work as a sequential, batch-1 Python loop (docs sahi/predict.py:270-320) with two host
syncs per detection (utils/yolo_wrapper.py:132,137).  Nothing here computes on the CPU: torch supplies device
memory, streams and the conv backbone only.
"""
from __future__ import annotations

import contextlib
import copy
import os

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _cabi, ops

ROW = ops.ROW
MAX_STAGE2_SEGMENT = 32768  # fsd_merge's max_segment limit


@dataclass
class SlicePlan:
    H: int
    W: int
    boxes: List[List[int]]          # [[x0,y0,x1,y1]] row-major, sahi semantics
    box_w: int
    box_h: int
    g_slice: dict                   # letterbox geometry of one slice
    g_full: Optional[dict]          # letterbox geometry of the full-image pass (None when it is skipped)

    @property
    def S(self) -> int:
        return len(self.boxes)


@dataclass
class DetectionBatch:
    """Host-side result of one batch: rows of image i are rows[offsets[i]:offsets[i+1]]."""
    offsets: np.ndarray             # [N+1]
    boxes: np.ndarray               # [T,4] float32 (integral values when the plugin truncates)
    scores: np.ndarray              # [T]
    keypoints: np.ndarray           # [T,5,3]
    has_keypoints: np.ndarray       # [T] bool
    stage1: Optional[Dict] = None   # per-slice detections (only on request)
    counters: Dict = field(default_factory=dict)

    def image(self, i: int):
        a, b = int(self.offsets[i]), int(self.offsets[i + 1])
        return self.boxes[a:b], self.scores[a:b], self.keypoints[a:b], self.has_keypoints[a:b]


def _box_pad(g: dict, src_w: int, src_h: int):
    """ultralytics scale_boxes pad (rounded) and scale_coords pad (not rounded) for one letterbox geometry."""
    gain = g["gain"]
    bx = round((g["out_w"] - src_w * gain) / 2 - 0.1)
    by = round((g["out_h"] - src_h * gain) / 2 - 0.1)
    return bx, by, (g["out_w"] - src_w * gain) / 2, (g["out_h"] - src_h * gain) / 2


class SlicedFaceDetector:
    def __init__(self, backbone: torch.nn.Module, device="cuda:0", imgsz: int = 1024, conf: float = 0.5,
                 half: bool = True, stride: int = 32, iou: float = 0.7, max_det: int = 300,
                 cap_per_entry: int = 1024, channels_last: bool = True, reverse_channels: bool = True,
                 chunk_entries: int = 96, truncate: bool = True, use_graphs: bool = False, overlap_post: bool = False):
        if not torch.cuda.is_available():
            raise _cabi.FsdError("SlicedFaceDetector needs a CUDA device: fsd_b200 has no CPU fallback")
        self.device = torch.device(device)
        self.dtype = torch.float16 if half else torch.float32
        self.imgsz, self.conf, self.stride, self.iou, self.max_det = imgsz, conf, stride, iou, max_det
        if os.environ.get("FSD_CHUNK_ENTRIES"):  # (measurement knob: network inputs per backbone call)
            chunk_entries = max(1, int(os.environ["FSD_CHUNK_ENTRIES"]))
        self.cap, self.reverse, self.chunk, self.truncate = cap_per_entry, reverse_channels, chunk_entries, truncate
        self.channels_last = channels_last
        # a private copy: Module.half() / .float() convert in place, and one YOLO front end may own an fp16 and an fp32 engine
        bb = copy.deepcopy(backbone).to(self.device).eval()
        bb = bb.half() if half else bb.float()
        for p in bb.parameters():
            p.requires_grad_(False)
        if channels_last:
            bb = bb.to(memory_format=torch.channels_last)
        self.backbone = bb
        # per-shape cuDNN algorithm search (the batched shapes repeat for every step) is switched on around the backbone
        # calls only (_cudnn_scope): constructing a detector does not change the process-wide torch.backends.cudnn flags
        self.cudnn_benchmark = True
        self._cudnn_limit = (int(os.environ["FSD_CUDNN_BENCHMARK_LIMIT"])  # 0 = try every algorithm (default: the first 10)
                             if os.environ.get("FSD_CUDNN_BENCHMARK_LIMIT") is not None else None)
        self._plans: Dict = {}
        self._dev_cache: Dict = {}
        self.handle = _cabi.get_handle(self.device.index or 0)
        self.head_hook = None  # tests: callable(entries_kind, x, levels) to record / replace head tensors
        # CUDA-graph replay of the backbone (opt-in): one graph per (stream, chunk address, shape) over static network-input
        # buffers, so a step enqueues a handful of graph launches instead of ~1300 kernel launches from Python
        self.use_graphs = use_graphs
        # synchronous detect(): run the slices' stage-1 NMS / finalize on a side stream under the full-image pass
        self.overlap_post = overlap_post
        self._own_post_stream = None
        self._graphs: Dict = {}
        self._graph_pools: Dict = {}
        self._xbufs: Dict = {}
        self.last_stage1 = None
        self.tie_rule = "box_lex"  # stage 2: sahi 0.11.34's equal-score rule (SURVEY A.2.4 variant N); "index" = plain greedy order
        self.replayed_launches = 0  # fsd kernels executed through graph replays (the handle only counts direct launches)

    # ------------------------------------------------------------------------------------------ planning
    def plan(self, H, W, slice_h, slice_w, ov_h, ov_w, perform_standard_pred=True) -> SlicePlan:
        key = (H, W, slice_h, slice_w, float(ov_h), float(ov_w), bool(perform_standard_pred), self.imgsz)
        p = self._plans.get(key)
        if p is None:
            boxes = _cabi.slice_plan(H, W, slice_h, slice_w, ov_h, ov_w)
            bw, bh = boxes[0][2] - boxes[0][0], boxes[0][3] - boxes[0][1]
            g_s = _cabi.letterbox_geometry(bh, bw, self.imgsz, self.stride)
            g_f = _cabi.letterbox_geometry(H, W, self.imgsz, self.stride) if (len(boxes) > 1 and perform_standard_pred) else None
            p = SlicePlan(H, W, boxes, bw, bh, g_s, g_f)
            self._plans[key] = p
        return p

    def _device_tables(self, plan: SlicePlan, N: int):
        """Entry lists and geometry tables for N images of one plan; built once and kept on the device."""
        key = (id(plan), N)
        t = self._dev_cache.get(key)
        if t is not None:
            return t
        dev, S = self.device, plan.S
        i32 = dict(dtype=torch.int32, device=dev)
        ent_s = torch.tensor([[i, b[0], b[1]] for i in range(N) for b in plan.boxes], **i32)
        bx, by, kx, ky = _box_pad(plan.g_slice, plan.box_w, plan.box_h)
        geo_s = torch.tensor([[b[0], b[1], plan.box_w, plan.box_h, bx, by, plan.W, plan.H] for _ in range(N) for b in plan.boxes], **i32)
        fgeo_s = torch.tensor([[plan.g_slice["gain"], kx, ky, 0.0]] * (N * S), dtype=torch.float32, device=dev)
        t = dict(ent_s=ent_s, geo_s=geo_s, fgeo_s=fgeo_s,
                 grange_s=torch.tensor([[i * S, (i + 1) * S] for i in range(N)], **i32),
                 seg_s=torch.arange(N * S, **i32) * self.cap)
        # worst case (S+1)*max_det per image, bounded by Kernel 3's segment limit: images with >= 109 slices (e.g. 8000x6000
        # at 640/0.2 = 192 slices) are accepted, and only an image that really produces more than 32768 per-slice detections
        # is refused (check_det_overflow) — never truncated silently
        det_cap = min((S + (1 if plan.g_full else 0)) * self.max_det, MAX_STAGE2_SEGMENT)
        t["det_cap"] = det_cap
        t["goff"] = torch.arange(N, **i32) * det_cap
        if plan.g_full is not None:
            bx, by, kx, ky = _box_pad(plan.g_full, plan.W, plan.H)
            t["ent_f"] = torch.tensor([[i, 0, 0] for i in range(N)], **i32)
            t["geo_f"] = torch.tensor([[0, 0, plan.W, plan.H, bx, by, plan.W, plan.H]] * N, **i32)
            t["fgeo_f"] = torch.tensor([[plan.g_full["gain"], kx, ky, 0.0]] * N, dtype=torch.float32, device=dev)
            t["grange_f"] = torch.tensor([[i, i + 1] for i in range(N)], **i32)
            t["seg_f"] = torch.arange(N, **i32) * self.cap
        self._dev_cache[key] = t
        return t

    @staticmethod
    def check_det_overflow(dmax: int, det_cap: int):
        if dmax > det_cap:
            raise _cabi.FsdError(f"an image produced {dmax} per-slice detections, more than the {det_cap} one stage-2 merge "
                                 "segment holds (Kernel 3 limit 32768): raise the confidence threshold or use merge_buffer_length")

    # ------------------------------------------------------------------------------------------ stages
    def _cudnn_scope(self):
        """cudnn.benchmark (and the optional search limit) for the backbone calls of this engine only."""
        return ops.cudnn_benchmark(self.cudnn_benchmark, self._cudnn_limit, tf32=None if self.dtype == torch.float16 else False)

    def _forward_entries(self, kind: str, x: torch.Tensor, cand: torch.Tensor, count: torch.Tensor):
        """backbone + Kernel 2a over a batch of network inputs, in chunks that bound activation memory."""
        E = x.shape[0]
        graphs = self.use_graphs and self.head_hook is None
        for a in range(0, E, self.chunk):
            xb = x[a:a + self.chunk]  # Kernel 1 already wrote the layout the backbone wants: no conversion pass
            with self._cudnn_scope():
                levels = self._graph_forward(xb) if graphs else self.backbone(xb)
            if self.head_hook is not None:
                levels = self.head_hook(kind, a, xb, levels)
            ops.pose_decode(levels, self.conf, cand=cand[a:a + self.chunk], count=count[a:a + self.chunk])

    def _network_input(self, kind: str, shape):
        """Static network-input buffer per (stream, kind, shape) when graphs are on (a captured graph reads fixed addresses)."""
        if not self.use_graphs:
            return None
        key = (torch.cuda.current_stream(self.device).cuda_stream, kind, tuple(shape))
        buf = self._xbufs.get(key)
        if buf is None:
            if len(self._xbufs) >= 16:
                return None  # shapes keep changing: stay eager
            buf = torch.empty(shape, dtype=self.dtype, device=self.device,
                              memory_format=torch.channels_last if self.channels_last else torch.contiguous_format)
            self._xbufs[key] = buf
        return buf

    def _graph_forward(self, xb: torch.Tensor):
        stream = torch.cuda.current_stream(self.device)
        key = (stream.cuda_stream, xb.data_ptr(), tuple(xb.shape))
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= 64 or not any(xb.data_ptr() >= b.data_ptr() and xb.data_ptr() < b.data_ptr() + b.numel() * b.element_size()
                                                   for b in self._xbufs.values()):
                return self.backbone(xb)  # not one of the static buffers: a graph would read a stale address
            self.backbone(xb)  # eager warm-up on this stream: cuDNN algorithm search must not run under capture
            pool = self._graph_pools.setdefault(stream.cuda_stream, torch.cuda.graph_pool_handle())
            graph = torch.cuda.CUDAGraph()
            n0 = self.handle.launches
            with torch.cuda.graph(graph, pool=pool):
                levels = self.backbone(xb)
            ent = (graph, levels, self.handle.launches - n0, xb)
            self._graphs[key] = ent
        ent[0].replay()
        self.replayed_launches += ent[2]
        return ent[1]

    @torch.no_grad()
    def measure_backbone_ms(self, plan: SlicePlan, N: int, steps: int = 3) -> float:
        """Device milliseconds per batch spent in the PyTorch backbone ALONE (all slice chunks + the full-image chunk, graph
        replays when enabled) over whatever the network-input buffers hold — measurement only: bench.py uses it for the
        backbone's share of the step instead of quoting a profile."""
        shapes = [("slices", (N * plan.S, 3, plan.g_slice["out_h"], plan.g_slice["out_w"]))]
        if plan.g_full is not None:
            shapes.append(("full", (N, 3, plan.g_full["out_h"], plan.g_full["out_w"])))
        bufs = []
        for kind, shape in shapes:
            x = self._network_input(kind, shape)
            if x is None:
                x = torch.zeros(shape, dtype=self.dtype, device=self.device).contiguous(
                    memory_format=torch.channels_last if self.channels_last else torch.contiguous_format)
            bufs.append(x)

        def once():
            for x in bufs:
                for a in range(0, x.shape[0], self.chunk):
                    with self._cudnn_scope():
                        self._graph_forward(x[a:a + self.chunk]) if self.use_graphs else self.backbone(x[a:a + self.chunk])

        once()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            once()
        t1.record()
        t1.synchronize()
        return t0.elapsed_time(t1) / steps

    def _stage1(self, cand, count, seg_off):
        return ops.merge_segments(cand.view(-1, ROW), seg_off, count, self.cap, merge_type="NMS", metric="IOU",
                                  thr=self.iou, cmp_strict=True, precision="fp32", class_agnostic=True,
                                  pre_cap=30000, max_keep=self.max_det, tie_col=5, want_parent=False)

    @torch.no_grad()
    def detect(self, pool: ops.ImagePool, slice_h: int, slice_w: int, ov_h: float = 0.2, ov_w: float = 0.2,
               perform_standard_pred: bool = True, postprocess_type: str = "GREEDYNMM", match_metric: str = "IOS",
               match_threshold: float = 0.5, class_agnostic: bool = False, want_stage1: bool = False,
               to_host: bool = True, post_stream: Optional[torch.cuda.Stream] = None):
        """Sliced detection of every image in `pool`; returns a DetectionBatch (or the device tensors if not to_host).

        With `post_stream` (only with to_host=False) everything after the network — stage-1 NMS, finalize, stage-2 merge,
        key-point attach, packing: latency-bound kernels on 32-384 CTAs — is enqueued on that stream behind events, so it
        overlaps the backbone of the full-image pass and of the NEXT batch on the calling stream.  The returned tensors
        then belong to `post_stream` (the dict carries it under "stream")."""
        if postprocess_type not in ("NMS", "GREEDYNMM", "NMM"):
            raise ValueError(f"postprocess_type should be one of ['GREEDYNMM', 'NMM', 'NMS', 'LSNMS'] but given as {postprocess_type}")
        N, dev = pool.n, self.device
        plan = self.plan(pool.h, pool.w, slice_h, slice_w, ov_h, ov_w, perform_standard_pred)
        t = self._device_tables(plan, N)
        S, E = plan.S, N * plan.S
        while True:  # retried with a larger candidate capacity if a slice overflowed it
            cand_s = torch.empty((E, self.cap, ROW), dtype=torch.float32, device=dev)
            count_s = torch.empty((E,), dtype=torch.int32, device=dev)
            x_s = ops.gather_letterbox(pool, t["ent_s"], plan.box_w, plan.box_h, self.imgsz, self.stride, self.reverse, self.dtype,
                                       channels_last=self.channels_last,
                                       out=self._network_input("slices", (E, 3, plan.g_slice["out_h"], plan.g_slice["out_w"])))
            self._forward_entries("slices", x_s, cand_s, count_s)
            del x_s
            main = torch.cuda.current_stream(dev)
            post = post_stream if not to_host else None
            if post is None and self.overlap_post:
                if self._own_post_stream is None:
                    self._own_post_stream = torch.cuda.Stream(device=dev)
                post = self._own_post_stream

            def on_post(*tensors):
                """Switch to the post stream behind everything enqueued so far; the caching allocator must not recycle
                `tensors` (allocated on the main stream) until the post stream is done with them."""
                if post is None:
                    return contextlib.nullcontext()
                ev = torch.cuda.Event()
                ev.record(main)
                post.wait_event(ev)
                for x in tensors:
                    x.record_stream(post)
                return torch.cuda.stream(post)

            with on_post(cand_s, count_s):
                s1 = self._stage1(cand_s, count_s, t["seg_s"])
                det = torch.empty((N * t["det_cap"], ROW), dtype=torch.float32, device=dev)
                dcount = torch.zeros((N,), dtype=torch.int32, device=dev)
                ops.finalize_dets(cand_s, s1["keep"], s1["keep_count"], t["geo_s"], t["fgeo_s"], t["grange_s"], t["goff"],
                                  det, dcount, t["det_cap"], self.truncate)
            count_f = None
            cand_f = None
            if plan.g_full is not None:
                cand_f = torch.empty((N, self.cap, ROW), dtype=torch.float32, device=dev)
                count_f = torch.empty((N,), dtype=torch.int32, device=dev)
                x_f = ops.gather_letterbox(pool, t["ent_f"], plan.W, plan.H, self.imgsz, self.stride, self.reverse, self.dtype,
                                           channels_last=self.channels_last,
                                           out=self._network_input("full", (N, 3, plan.g_full["out_h"], plan.g_full["out_w"])))
                self._forward_entries("full", x_f, cand_f, count_f)
                del x_f
            with on_post(*([cand_f, count_f] if cand_f is not None else [])):
                if cand_f is not None:
                    s1f = self._stage1(cand_f, count_f, t["seg_f"])
                    ops.finalize_dets(cand_f, s1f["keep"], s1f["keep_count"], t["geo_f"], t["fgeo_f"], t["grange_f"],
                                      t["goff"], det, dcount, t["det_cap"], self.truncate)
                # stage 2: cross-slice merge per image (skipped by the reference when an image has <= 1 prediction:
                # a single box is its own keep, so running it is equivalent)
                # (this path has ONE category, "face": without a category array Kernel 3 treats every pair as same-class, so
                # class_agnostic True and False give the same result — exactly as sahi's batched_* variants do for one class)
                s2 = ops.merge_segments(det, t["goff"], dcount, t["det_cap"], merge_type=postprocess_type,
                                        metric=match_metric, thr=match_threshold, cmp_strict=False, precision="fp64",
                                        class_agnostic=bool(class_agnostic), want_parent=False, tie_rule=self.tie_rule)
                # per-image stage-1 detections BEFORE the det_cap clamp (overflow is reported, never silently truncated)
                dmax = s1["keep_count"].view(N, S).sum(1)
                if cand_f is not None:
                    dmax = dmax + s1f["keep_count"]
                dmax = dmax.max()
                src = ops.attach_keypoints(s2["boxes"], t["goff"], s2["keep_count"], det, t["goff"], dcount)
                rows, offsets = ops.pack_results(det, t["goff"], s2, src)
            if post is not None and not to_host:
                return dict(rows=rows, offsets=offsets, det=det, dcount=dcount, count_s=count_s, count_f=count_f, plan=plan,
                            stream=post, cap=self.cap, tables=t, dmax=dmax)
            if post is not None:
                main.wait_stream(post)  # synchronous form: the results are read on the calling stream below
            if not to_host:
                return dict(rows=rows, offsets=offsets, det=det, dcount=dcount, count_s=count_s, count_f=count_f, plan=plan,
                            cap=self.cap, tables=t, dmax=dmax)
            # ---- the only device->host traffic of the batch: counts, then exactly the packed rows
            off_h = offsets.cpu().numpy()
            cmax = int(count_s.max()) if count_f is None else max(int(count_s.max()), int(count_f.max()))
            self.check_det_overflow(int(dmax), t["det_cap"])
            if cmax > self.cap:
                self.cap = 1 << (cmax - 1).bit_length()
                self._dev_cache.clear()
                t = self._device_tables(plan, N)
                continue
            break
        total = int(off_h[-1])
        host = rows[:total].cpu().numpy()
        srcs = host[:, 5].view(np.int32) if total else np.zeros((0,), np.int32)
        out = DetectionBatch(offsets=off_h, boxes=host[:, :4].copy(), scores=host[:, 4].copy(),
                             keypoints=host[:, 6:21].reshape(-1, 5, 3).copy(), has_keypoints=srcs >= 0,
                             counters=dict(slices=S, entries=E + (N if plan.g_full else 0), cap=self.cap))
        if want_stage1:
            dc = dcount.cpu().numpy()
            d = det.cpu().numpy().reshape(N, t["det_cap"], ROW)
            out.stage1 = dict(count=dc, rows=[d[i, : dc[i]].copy() for i in range(N)])
            self.last_stage1 = out.stage1  # parity tests compare the per-slice detections behind the reference-facing call
        return out

    # ------------------------------------------------------------------------------------------ single entry
    @torch.no_grad()
    def predict_array(self, img: np.ndarray, conf: Optional[float] = None, imgsz: Optional[int] = None):
        """ultralytics-style single prediction on one HWC uint8 array: returns (boxes [n,4] f32 un-letterboxed,
        scores [n], keypoints [n,5,3]) as CUDA tensors, score-descending (YOLO.predict surface)."""
        old = (self.conf, self.imgsz)
        if conf is not None:
            self.conf = conf
        if imgsz is not None:
            self.imgsz = imgsz
        try:
            H, W = img.shape[:2]
            pool = ops.ImagePool.from_numpy([img], self.device)
            g = _cabi.letterbox_geometry(H, W, self.imgsz, self.stride)
            bx, by, kx, ky = _box_pad(g, W, H)
            dev = self.device
            i32 = dict(dtype=torch.int32, device=dev)
            while True:
                cand = torch.empty((1, self.cap, ROW), dtype=torch.float32, device=dev)
                count = torch.empty((1,), dtype=torch.int32, device=dev)
                x = ops.gather_letterbox(pool, torch.zeros((1, 3), **i32), W, H, self.imgsz, self.stride, self.reverse, self.dtype,
                                         channels_last=self.channels_last)
                self._forward_entries("single", x, cand, count)
                if int(count[0]) > self.cap:
                    self.cap = 1 << (int(count[0]) - 1).bit_length()
                    self._dev_cache.clear()
                    continue
                break
            s1 = self._stage1(cand, count, torch.zeros((1,), **i32))
            det = torch.empty((self.max_det, ROW), dtype=torch.float32, device=dev)
            dcount = torch.zeros((1,), **i32)
            ops.finalize_dets(cand, s1["keep"], s1["keep_count"], torch.tensor([[0, 0, W, H, bx, by, 0, 0]], **i32),
                              torch.tensor([[g["gain"], kx, ky, 0.0]], dtype=torch.float32, device=dev),
                              torch.tensor([[0, 1]], **i32), torch.zeros((1,), **i32), det, dcount, self.max_det,
                              truncate=False)
            n = int(dcount[0])
            return det[:n, :4], det[:n, 4], det[:n, 6:21].reshape(n, 5, 3)
        finally:
            self.conf, self.imgsz = old
