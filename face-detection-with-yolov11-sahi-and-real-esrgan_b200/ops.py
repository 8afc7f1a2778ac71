"""Torch-tensor front ends of the C-ABI kernels.

PyTorch is used here for device memory, streams and dtype bookkeeping only; every byte of arithmetic on the hot
path is done by the hand-written sm_100a kernels in csrc/.  All functions raise if CUDA is unavailable.
"""
from __future__ import annotations

import os

import contextlib
import ctypes as C
import warnings

import numpy as np
import torch

from . import _cabi
from ._cabi import FSD_F16, FSD_F32, check, get_handle

_TORCH_DTYPE = {torch.float16: FSD_F16, torch.float32: FSD_F32}


@contextlib.contextmanager
def cudnn_benchmark(on: bool = True, limit: int | None = None, tf32: bool | None = None):
    """Per-shape cuDNN algorithm search around the library convolutions of the backbones, restored on exit: the product
    never leaves process-wide torch.backends flags changed behind the caller's back.  `tf32=False` additionally forces true
    fp32 convolutions / matmuls inside the scope (the fp32 engine: the reference's CPU path computes in IEEE fp32, and
    TF32's 10-bit mantissa is fp16-grade noise)."""
    cud, mm = torch.backends.cudnn, torch.backends.cuda.matmul
    old = (cud.benchmark, cud.benchmark_limit, cud.allow_tf32, mm.allow_tf32)
    cud.benchmark = bool(on)
    if limit is not None:
        cud.benchmark_limit = int(limit)
    if tf32 is not None:
        cud.allow_tf32 = mm.allow_tf32 = bool(tf32)
    try:
        yield
    finally:
        cud.benchmark, cud.benchmark_limit, cud.allow_tf32, mm.allow_tf32 = old


_WARNED_CPU = []


def resolve_device(device) -> str:
    """The CUDA device a plug-in runs on.  The reference's plug-ins default to device='cpu' (utils/yolo_wrapper.py:8); this
    library has no CPU path, so a CPU request is redirected to cuda:0 with ONE warning per process (never silently), and
    raises when no CUDA device is visible."""
    d = str(device)
    if d in ("cpu", "None", ""):
        if not torch.cuda.is_available():
            raise _cabi.FsdError("fsd_b200 runs on CUDA only (no CPU fallback) and no CUDA device is visible")
        if not _WARNED_CPU:
            _WARNED_CPU.append(d)
            warnings.warn(f"device={device!r} requested, but fsd_b200 has no CPU path: running on cuda:0", RuntimeWarning, stacklevel=3)
        return "cuda:0"
    return d


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _cabi.FsdError(f"{name} must be a CUDA tensor: fsd_b200 has no CPU fallback")


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def _handle_for(t: torch.Tensor):
    return get_handle(t.device.index if t.device.index is not None else torch.cuda.current_device())


class ImagePool:
    """N same-sized HWC uint8 images resident in HBM as one [N, H, pitch] buffer (pitch = W*3 rounded up to 16 B).

    This is the layout Kernel 1's TMA tensor map walks: dims {pitch/4 u32, H, N}.
    """

    def __init__(self, n: int, height: int, width: int, device="cuda"):
        self.n, self.h, self.w = int(n), int(height), int(width)
        self.pitch = (self.w * 3 + 15) // 16 * 16
        self.buf = torch.zeros((self.n, self.h, self.pitch), dtype=torch.uint8, device=device)

    @property
    def device(self):
        return self.buf.device

    @property
    def image_pitch(self) -> int:
        return self.h * self.pitch

    def upload(self, index: int, img, non_blocking: bool = True):
        """Copy one HWC uint8 image (numpy array or CPU/pinned tensor) into slot `index`."""
        t = torch.from_numpy(np.ascontiguousarray(img)) if isinstance(img, np.ndarray) else img
        if t.shape != (self.h, self.w, 3) or t.dtype != torch.uint8:
            raise ValueError(f"expected uint8 [{self.h},{self.w},3], got {tuple(t.shape)} {t.dtype}")
        self.buf[index, :, : self.w * 3].copy_(t.reshape(self.h, self.w * 3), non_blocking=non_blocking)

    def upload_jpeg(self, index: int, data: bytes, bgr: bool = False):
        """(f4) decode a JPEG (bytes) with nvJPEG straight into slot `index` — RGB like PIL's convert("RGB"), or BGR like
        cv2.imread.  Only the compressed bytes cross PCIe.  Pixels differ from PIL's libjpeg by a few LSB: opt-in."""
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        h = get_handle(self.buf.device.index if self.buf.device.index is not None else torch.cuda.current_device())
        slot = self.buf[index]
        check(h.lib.fsd_jpeg_decode(h.h, buf, len(data), 1 if bgr else 0, slot.data_ptr(), self.pitch, self.h, self.w,
                                    _stream_ptr(self.buf.device)), "fsd_jpeg_decode")
        torch.cuda.current_stream(self.buf.device).synchronize()  # nvJPEG reads the host bytes asynchronously: keep `buf` alive until done

    def view(self, index: int) -> torch.Tensor:
        """[H, W, 3] view of slot `index` (strided when the pitch is padded)."""
        return self.buf[index, :, : self.w * 3].unflatten(1, (self.w, 3))

    def subpool(self, start: int, stop: int) -> "ImagePool":
        """View of images [start, stop) sharing this pool's memory."""
        sub = object.__new__(ImagePool)
        sub.n, sub.h, sub.w, sub.pitch = stop - start, self.h, self.w, self.pitch
        sub.buf = self.buf[start:stop]
        return sub

    @classmethod
    def from_numpy(cls, images, device="cuda") -> "ImagePool":
        h, w = images[0].shape[:2]
        pool = cls(len(images), h, w, device)
        for i, im in enumerate(images):
            pool.upload(i, im, non_blocking=False)
        return pool


def gather_letterbox(pool: ImagePool, entries: torch.Tensor, src_w: int, src_h: int, imgsz: int = 1024,
                     stride: int = 32, reverse_channels: bool = True, dtype=torch.float16,
                     out: torch.Tensor | None = None, channels_last: bool = False) -> torch.Tensor:
    """Kernel 1: entries int32 [B,3] (image_index, x0, y0) -> [B,3,out_h,out_w] network input (NCHW or channels-last)."""
    _require_cuda(pool.buf, "image pool")
    g = _cabi.letterbox_geometry(src_h, src_w, imgsz, stride)
    B = int(entries.shape[0])
    if out is None:
        out = torch.empty((B, 3, g["out_h"], g["out_w"]), dtype=dtype, device=pool.device,
                          memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    if out.is_contiguous():
        layout = _cabi.FSD_PLANAR
    elif out.is_contiguous(memory_format=torch.channels_last):
        layout = _cabi.FSD_CHANNELS_LAST
    else:
        raise ValueError("gather_letterbox output must be dense NCHW or channels_last")
    if entries.dtype != torch.int32 or not entries.is_cuda or not entries.is_contiguous():
        entries = entries.to(device=pool.device, dtype=torch.int32).contiguous()
    h = _handle_for(pool.buf)
    check(h.lib.fsd_gather_letterbox(h.h, pool.buf.data_ptr(), pool.n, pool.h, pool.w, pool.pitch,
                                     pool.image_pitch, entries.data_ptr(), B, src_w, src_h, imgsz, stride,
                                     1 if reverse_channels else 0, _TORCH_DTYPE[out.dtype], layout, out.data_ptr(),
                                     _stream_ptr(pool.device)), "fsd_gather_letterbox")
    return out


# ---- Kernel 2a ------------------------------------------------------------------------------------------
ROW = 24  # floats per candidate / detection row (see include/fsd_b200.h)


def _level_layout(levels):
    """(layout, dtype) of the head tensors; all nine must agree and be dense in NCHW or channels-last order."""
    t0 = levels[0][0]
    layout = None
    for lvl in levels:
        for t in lvl:
            _require_cuda(t, "head tensor")
            if t.dtype != t0.dtype:
                raise ValueError("head tensors must share one dtype")
            if t.is_contiguous():
                lay = _cabi.FSD_PLANAR
            elif t.is_contiguous(memory_format=torch.channels_last):
                lay = _cabi.FSD_CHANNELS_LAST
            else:
                raise ValueError("head tensors must be contiguous (NCHW) or channels_last")
            if t.shape[1] == 1 or t.shape[2] * t.shape[3] == 1:
                continue  # ambiguous strides: a 1-channel tensor is both layouts at once
            if layout is None:
                layout = lay
            elif layout != lay:
                raise ValueError("head tensors mix NCHW and channels_last")
    return (layout if layout is not None else _cabi.FSD_PLANAR), _TORCH_DTYPE[t0.dtype]


def pose_decode(levels, conf: float, cap_per_entry: int = 1024, cand: torch.Tensor | None = None,
                count: torch.Tensor | None = None):
    """Kernel 2a.  levels = [(box [B,64,h,w], cls [B,1,h,w], kpt [B,15,h,w])] for strides 8/16/32.
    Returns (cand [B, cap, 24] f32, count [B] i32)."""
    layout, dt = _level_layout(levels)
    B = int(levels[0][0].shape[0])
    dev = levels[0][0].device
    if levels[0][1].shape[1] != 1 or levels[0][2].shape[1] != 15 or levels[0][0].shape[1] != 64:
        raise ValueError("pose_decode expects nc=1, kpt_shape=(5,3), reg_max=16 head tensors")
    if cand is None:
        cand = torch.empty((B, cap_per_entry, ROW), dtype=torch.float32, device=dev)
    if count is None:
        count = torch.empty((B,), dtype=torch.int32, device=dev)
    arr = lambda k: (C.c_void_p * 3)(*[lvl[k].data_ptr() for lvl in levels])  # noqa: E731
    hw = (C.c_int32 * 6)(*[int(v) for lvl in levels for v in lvl[0].shape[2:]])
    h = _handle_for(cand)
    check(h.lib.fsd_pose_decode(h.h, arr(0), arr(1), arr(2), hw, B, layout, dt, float(conf), cand.data_ptr(),
                                int(cand.shape[1]), count.data_ptr(), _stream_ptr(dev)), "fsd_pose_decode")
    return cand, count


# ---- Kernel 3 -------------------------------------------------------------------------------------------
_MERGE_TYPE = {"NMS": _cabi.FSD_NMS, "GREEDYNMM": _cabi.FSD_GREEDYNMM, "NMM": _cabi.FSD_NMM}
_METRIC = {"IOU": _cabi.FSD_IOU, "IOS": _cabi.FSD_IOS}
_TIE_RULE = {"index": 0, "box_lex": 1}  # include/fsd_b200.h: fsd_merge tie_rule


def merge_segments(rows: torch.Tensor, seg_offsets: torch.Tensor, seg_counts: torch.Tensor | None, max_segment: int,
                   merge_type="NMS", metric="IOU", thr=0.5, cmp_strict=False, precision="fp64", class_agnostic=True,
                   pre_cap=0, max_keep=0, box_col=0, score_col=4, tie_col=None, cats: torch.Tensor | None = None,
                   want_parent=True, tie_rule="index"):
    """Kernel 3 over a [N, R] float32 row matrix (boxes at box_col..+3, score at score_col, optional int-bits
    tie-break key at tie_col).  tie_rule "box_lex" = sahi 0.11.34's equal-score rule (stage 2), "index" = the plain greedy
    loop (torchvision / stage 1).  Returns dict(keep, keep_count, parent, boxes, scores, cats)."""
    _require_cuda(rows, "rows")
    assert rows.dtype == torch.float32 and rows.dim() == 2 and rows.is_contiguous()
    N, R = rows.shape
    dev = rows.device
    S = int(seg_offsets.shape[0])
    keep = torch.empty((N,), dtype=torch.int32, device=dev)
    keep_count = torch.empty((S,), dtype=torch.int32, device=dev)
    parent = torch.full((N,), -1, dtype=torch.int32, device=dev) if want_parent else None
    mboxes = torch.empty((N, 4), dtype=torch.float32, device=dev)
    mscores = torch.empty((N,), dtype=torch.float32, device=dev)
    mcats = torch.empty((N,), dtype=torch.int32, device=dev) if cats is not None else None
    h = _handle_for(rows)
    ws_bytes = int(h.lib.fsd_merge_workspace_bytes(N, S, int(max_segment)))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    base = rows.data_ptr()
    check(h.lib.fsd_merge(h.h, base + 4 * box_col, R, base + 4 * score_col, R,
                          _ptr(cats), 1, (base + 4 * tie_col) if tie_col is not None else 0, R,
                          seg_offsets.data_ptr(), _ptr(seg_counts), S, int(max_segment),
                          _MERGE_TYPE[merge_type], _METRIC[metric], float(thr), 1 if cmp_strict else 0,
                          0 if precision == "fp64" else 1, 1 if class_agnostic else 0, int(pre_cap), int(max_keep),
                          _TIE_RULE[tie_rule], keep.data_ptr(), keep_count.data_ptr(), _ptr(parent), mboxes.data_ptr(),
                          mscores.data_ptr(), _ptr(mcats), ws.data_ptr(), ws_bytes, _stream_ptr(dev)), "fsd_merge")
    return dict(keep=keep, keep_count=keep_count, parent=parent, boxes=mboxes, scores=mscores, cats=mcats)


# ---- Kernel 2b ------------------------------------------------------------------------------------------
def finalize_dets(cand, keep, keep_count, entry_geom, entry_fgeom, group_range, group_offsets, det, out_count,
                  det_cap_per_group: int, truncate: bool = True):
    """Kernel 2b: append the kept rows of every entry to its image's detection list (see the header)."""
    B, cap = int(cand.shape[0]), int(cand.shape[1])
    G = int(group_range.shape[0])
    h = _handle_for(cand)
    check(h.lib.fsd_finalize_dets(h.h, cand.data_ptr(), cap, keep.data_ptr(), keep_count.data_ptr(), B,
                                  entry_geom.data_ptr(), entry_fgeom.data_ptr(), group_range.data_ptr(),
                                  group_offsets.data_ptr(), G, 1 if truncate else 0, det.data_ptr(),
                                  int(det_cap_per_group), out_count.data_ptr(), _stream_ptr(cand.device)),
          "fsd_finalize_dets")
    return det, out_count


def pack_results(det, group_offsets, s2, src_index, out: torch.Tensor | None = None):
    """(a1) pack stage-2 results of all images into contiguous rows; returns (rows [cap,24], offsets [G+1] i32)."""
    G = int(group_offsets.shape[0])
    dev = det.device
    if out is None:
        out = torch.empty((det.shape[0], ROW), dtype=torch.float32, device=dev)
    offsets = torch.empty((G + 1,), dtype=torch.int32, device=dev)
    h = _handle_for(det)
    check(h.lib.fsd_pack_results(h.h, det.data_ptr(), group_offsets.data_ptr(), s2["keep"].data_ptr(),
                                 s2["keep_count"].data_ptr(), s2["boxes"].data_ptr(), s2["scores"].data_ptr(),
                                 _ptr(src_index), G, out.data_ptr(), offsets.data_ptr(), _stream_ptr(dev)),
          "fsd_pack_results")
    return out, offsets


_ACT = {None: 0, "none": 0, "silu": 1, "lrelu": 2}


def bias_act_(x: torch.Tensor, bias: torch.Tensor, act: str = "silu", slope: float = 0.2) -> torch.Tensor:
    """(a5) in-place conv epilogue x = act(x + bias[c]) for a channels-last CUDA tensor [N,C,H,W] (one HBM pass)."""
    _require_cuda(x, "x")
    n, c, hh, ww = x.shape
    if not x.is_contiguous(memory_format=torch.channels_last):
        raise ValueError("bias_act_ needs a dense channels_last tensor")
    h = _handle_for(x)
    check(h.lib.fsd_bias_act_inplace(h.h, x.data_ptr(), bias.data_ptr(), n * hh * ww, c, _ACT[act], float(slope),
                                     _TORCH_DTYPE[x.dtype], _stream_ptr(x.device)), "fsd_bias_act_inplace")
    return x


def _slot_stride(t: torch.Tensor, n: int, c: int, hh: int, ww: int, name: str) -> int:
    """Pixel stride (elements) of `t`, a [N,C,H,W] channel slot of a dense channels-last buffer."""
    if tuple(t.shape) != (n, c, hh, ww):
        raise ValueError(f"{name}: expected shape {(n, c, hh, ww)}, got {tuple(t.shape)}")
    ct = t.stride(3) if ww > 1 else (t.stride(2) if hh > 1 else t.stride(0))
    ok = (c == 1 or t.stride(1) == 1) and (ww == 1 or t.stride(3) == ct) and (hh == 1 or t.stride(2) == ww * ct) \
        and (n == 1 or t.stride(0) == hh * ww * ct) and ct >= c
    if not ok:
        raise ValueError(f"{name} must be a channel slot of a dense channels_last tensor (strides {t.stride()})")
    return int(ct)


def bias_act(x: torch.Tensor, bias: torch.Tensor, act: str = "silu", slope: float = 0.2, out: torch.Tensor | None = None,
             residual: torch.Tensor | None = None, out2: torch.Tensor | None = None,
             up2: torch.Tensor | None = None) -> torch.Tensor:
    """(a5) general fp16 conv epilogue: out = act(x + bias[c]) (+ residual).  `x` is the dense channels-last raw
    convolution [N,C,H,W]; `out` (default: x itself), `residual` and `out2` may be channel slots (views `buf[:, a:b]`) of
    wider channels-last buffers.  `out2` [N,C2,H,W] additionally receives the LAST C2 channels; `up2` [N,C,2H,2W] (a slot
    too) receives the result up-sampled 2x (nearest).  Returns `out`."""
    _require_cuda(x, "x")
    n, c, hh, ww = x.shape
    if x.dtype != torch.float16 or not x.is_contiguous(memory_format=torch.channels_last):
        raise ValueError("bias_act needs a dense channels_last fp16 tensor")
    if out is None:
        out = x
    so = _slot_stride(out, n, c, hh, ww, "out")
    sr = _slot_stride(residual, n, c, hh, ww, "residual") if residual is not None else 0
    c2 = int(out2.shape[1]) if out2 is not None else 0
    s2 = _slot_stride(out2, n, c2, hh, ww, "out2") if out2 is not None else 0
    su = _slot_stride(up2, n, c, 2 * hh, 2 * ww, "up2") if up2 is not None else 0
    h = _handle_for(x)
    check(h.lib.fsd_bias_act(h.h, x.data_ptr(), bias.data_ptr(), out.data_ptr(), so,
                             residual.data_ptr() if residual is not None else None, sr,
                             out2.data_ptr() if out2 is not None else None, s2, c - c2 if out2 is not None else 0,
                             up2.data_ptr() if up2 is not None else None, su, hh, ww,
                             n * hh * ww, c, _ACT[act], float(slope), _TORCH_DTYPE[x.dtype], _stream_ptr(x.device)),
          "fsd_bias_act")
    return out


def stem_conv(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, out: torch.Tensor | None = None,
              space_to_depth: bool = False) -> torch.Tensor:
    """(a5) SiLU(conv2d(x, weight, bias, stride=2, padding=1)) for the 3-channel fp16 channels-last network input, one kernel.
    x [E,3,H,W] channels-last dense, weight [16,3,3,3] standard-contiguous, bias [16] -> [E,16,H/2,W/2] channels-last, or with
    `space_to_depth` -> [E,64,H/4+1,W/4+1] (zero first row / column, 2x2 blocks folded into channels: see fsd_stem_conv)."""
    _require_cuda(x, "x")
    e, c, hh, ww = x.shape
    if c != 3 or x.dtype != torch.float16 or not x.is_contiguous(memory_format=torch.channels_last) or ww % 2:
        raise ValueError("stem_conv needs a dense channels_last fp16 [E,3,H,W] input with even W")
    if tuple(weight.shape) != (16, 3, 3, 3) or not weight.is_contiguous() or weight.dtype != torch.float16:
        raise ValueError("stem_conv needs a contiguous fp16 [16,3,3,3] weight")
    oh, ow = (hh - 1) // 2 + 1, (ww - 1) // 2 + 1
    if space_to_depth:
        if oh % 2 or ow % 2:
            raise ValueError("stem_conv(space_to_depth=True) needs even output sizes")
        shape = (e, 64, oh // 2 + 1, ow // 2 + 1)
    else:
        shape = (e, 16, oh, ow)
    if out is None:
        out = torch.empty(shape, dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        if space_to_depth:  # the kernel never writes the padding row / column
            out[:, :, 0].zero_()
            out[:, :, :, 0].zero_()
    elif tuple(out.shape) != shape or not out.is_contiguous(memory_format=torch.channels_last):
        raise ValueError(f"stem_conv: out must be a dense channels_last {shape} tensor")
    h = _handle_for(x)
    check(h.lib.fsd_stem_conv(h.h, x.data_ptr(), e, hh, ww, weight.data_ptr(), bias.data_ptr(), 16, _TORCH_DTYPE[x.dtype],
                              1 if space_to_depth else 0, out.data_ptr(), _stream_ptr(x.device)), "fsd_stem_conv")
    return out


def pointwise_tc_enabled() -> bool:
    """The tensor-core 1x1 kernel (csrc/k10_pointwise_tc.cu) is on unless FSD_K7_NO_TC is set."""
    return not os.environ.get("FSD_K7_NO_TC")


def pointwise_tc_supported(k: int, n: int) -> bool:
    """Shapes the tcgen05 path of fsd_pointwise_conv takes: channels in multiples of 16, weights <= 96 KB (resident in smem)."""
    return int(_cabi.load_library().fsd_pointwise_conv_supported(int(k), int(n))) == 2  # (the library owns the rule)


def pointwise_tc_preferred(k: int, n: int) -> bool:
    """Where the tcgen05 kernel beats both alternatives (profiles/r2_kernels_k10.jsonl): everything it takes (N = 256 runs one CTA per
    SM with sixteen epilogue warps: 0.43 of peak against 0.35 for cuDNN + epilogue)."""
    return pointwise_tc_supported(k, n)


def pointwise_conv_supported(k: int, n: int) -> bool:
    """Layer shapes fsd_pointwise_conv takes (the rest stays on the library convolution + fsd_bias_act)."""
    level = int(_cabi.load_library().fsd_pointwise_conv_supported(int(k), int(n)))
    return level == 2 if pointwise_tc_enabled() and level == 2 else (k % 16 == 0 and 16 <= k <= 128 and n in (16, 32, 64, 128))


def pointwise_conv(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, act: str = "silu", slope: float = 0.2,
                   out: torch.Tensor | None = None, residual: torch.Tensor | None = None,
                   out2: torch.Tensor | None = None) -> torch.Tensor:
    """(a5) act(conv1x1(x) + bias) (+ residual) in one kernel.  x [B,K,H,W] and out [B,N,H,W] (default: a new dense
    tensor) / residual / out2 may be channel slots of channels-last fp16 buffers; weight [N,K] (or [N,K,1,1])."""
    _require_cuda(x, "x")
    b, k, hh, ww = x.shape
    n = int(weight.shape[0])
    if x.dtype != torch.float16:
        raise ValueError("pointwise_conv needs fp16 tensors")
    w2 = weight.reshape(n, k)
    if not w2.is_contiguous():
        w2 = w2.contiguous()
    sx = _slot_stride(x, b, k, hh, ww, "x")
    if out is None:
        out = torch.empty((b, n, hh, ww), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    so = _slot_stride(out, b, n, hh, ww, "out")
    sr = _slot_stride(residual, b, n, hh, ww, "residual") if residual is not None else 0
    c2 = int(out2.shape[1]) if out2 is not None else 0
    s2 = _slot_stride(out2, b, c2, hh, ww, "out2") if out2 is not None else 0
    h = _handle_for(x)
    check(h.lib.fsd_pointwise_conv(h.h, x.data_ptr(), sx, w2.data_ptr(), bias.data_ptr(), out.data_ptr(), so,
                                   residual.data_ptr() if residual is not None else None, sr,
                                   out2.data_ptr() if out2 is not None else None, s2, n - c2 if out2 is not None else 0,
                                   b * hh * ww, k, n, _ACT[act], float(slope), _TORCH_DTYPE[x.dtype], _stream_ptr(x.device)),
          "fsd_pointwise_conv")
    return out


def conv3x3_tc_enabled() -> bool:
    """The tensor-core 3x3 convolution (fsd_conv3x3) is on unless FSD_NO_CONV3_TC is set."""
    return not os.environ.get("FSD_NO_CONV3_TC")


def conv3x3_preferred(k: int, n: int) -> bool:
    """Where fsd_conv3x3 beats cuDNN + fsd_bias_act (profiles/r2_kernels_conv3.jsonl): one channel slab (K <= 64), i.e. the halo mode with
    ONE TMA box per tile — 10-45 % faster on the YOLO11n shapes; with K >= 96 (one box per tap and slab) the library pair wins."""
    return k <= 64 and conv3x3_supported(k, n)


def conv3x3_supported(k: int, n: int) -> bool:
    """Shapes fsd_conv3x3 takes (stride 1, pad 1, dense): asks the library, which owns the rule."""
    return bool(_cabi.load_library().fsd_conv3x3_supported(int(k), int(n)))


def conv3x3_tap_major(weight: torch.Tensor) -> torch.Tensor:
    """[N, K, 3, 3] convolution weight -> the tap-major [3, 3, n, K] fp16 matrix fsd_conv3x3 reads (n = 16 with zero rows 8..15 when N = 8)."""
    n, k = int(weight.shape[0]), int(weight.shape[1])
    w = weight.detach().permute(2, 3, 0, 1).contiguous()
    if n == 8:
        w = torch.cat([w, torch.zeros_like(w)], dim=2).contiguous()
    return w


def conv3x3(x: torch.Tensor, weight_taps: torch.Tensor, bias: torch.Tensor, act: str = "silu", slope: float = 0.2,
            out: torch.Tensor | None = None, residual: torch.Tensor | None = None) -> torch.Tensor:
    """(a5) act(conv3x3(x, stride 1, pad 1) + bias) (+ residual) in one tensor-core kernel.  x [B,K,H,W] and out [B,N,H,W] (default: a
    new dense tensor) / residual may be channel slots of channels-last fp16 buffers; weight_taps from conv3x3_tap_major()."""
    _require_cuda(x, "x")
    b, k, hh, ww = x.shape
    n = int(bias.shape[0])
    if x.dtype != torch.float16 or weight_taps.dtype != torch.float16:
        raise ValueError("conv3x3 needs fp16 tensors")
    if tuple(weight_taps.shape) != (3, 3, 16 if n == 8 else n, k) or not weight_taps.is_contiguous():
        raise ValueError("conv3x3: weight_taps must be the contiguous tap-major matrix of conv3x3_tap_major()")
    sx = _slot_stride(x, b, k, hh, ww, "x")
    if out is None:
        out = torch.empty((b, n, hh, ww), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    so = _slot_stride(out, b, n, hh, ww, "out")
    sr = _slot_stride(residual, b, n, hh, ww, "residual") if residual is not None else 0
    h = _handle_for(x)
    check(h.lib.fsd_conv3x3(h.h, x.data_ptr(), sx, b, hh, ww, weight_taps.data_ptr(), bias.data_ptr(), out.data_ptr(), so,
                            residual.data_ptr() if residual is not None else None, sr, k, n, _ACT[act], float(slope),
                            _TORCH_DTYPE[x.dtype], _stream_ptr(x.device)), "fsd_conv3x3")
    return out


def conv2x2_supported(k: int, n: int) -> bool:
    return bool(_cabi.load_library().fsd_conv2x2_supported(int(k), int(n)))


def conv2x2(x: torch.Tensor, weight_taps: torch.Tensor, bias: torch.Tensor, act: str = "silu", slope: float = 0.2,
            out: torch.Tensor | None = None) -> torch.Tensor:
    """(a5) act(conv2x2(x, stride 1, no padding) + bias) on the tensor cores: x [B,K,H,W] -> [B,N,H-1,W-1]; weight_taps is the
    contiguous tap-major [2, 2, N, K] fp16 matrix (`w.permute(2, 3, 0, 1).contiguous()`)."""
    _require_cuda(x, "x")
    b, k, hh, ww = x.shape
    n = int(bias.shape[0])
    if x.dtype != torch.float16 or tuple(weight_taps.shape) != (2, 2, n, k) or not weight_taps.is_contiguous() or weight_taps.dtype != torch.float16:
        raise ValueError("conv2x2 needs fp16 tensors and a contiguous tap-major [2, 2, N, K] weight")
    sx = _slot_stride(x, b, k, hh, ww, "x")
    if out is None:
        out = torch.empty((b, n, hh - 1, ww - 1), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    so = _slot_stride(out, b, n, hh - 1, ww - 1, "out")
    h = _handle_for(x)
    check(h.lib.fsd_conv2x2(h.h, x.data_ptr(), sx, b, hh, ww, weight_taps.data_ptr(), bias.data_ptr(), out.data_ptr(), so, k, n,
                            _ACT[act], float(slope), _TORCH_DTYPE[x.dtype], _stream_ptr(x.device)), "fsd_conv2x2")
    return out


def dwconv3x3_tap_major(weight: torch.Tensor) -> torch.Tensor:
    """[C, 1, 3, 3] depth-wise weight -> the tap-major [9, C] fp16 matrix fsd_dwconv3x3 reads."""
    c = int(weight.shape[0])
    return weight.detach().reshape(c, 9).t().contiguous()


def dwconv3x3(x: torch.Tensor, weight_taps: torch.Tensor, bias: torch.Tensor, act: str = "silu", slope: float = 0.2,
              out: torch.Tensor | None = None) -> torch.Tensor:
    """(a5) act(depthwise_conv3x3(x, stride 1, pad 1) + bias) in one kernel; x / out may be channel slots of channels-last fp16 buffers."""
    _require_cuda(x, "x")
    b, c, hh, ww = x.shape
    if x.dtype != torch.float16 or weight_taps.dtype != torch.float16 or tuple(weight_taps.shape) != (9, c) or not weight_taps.is_contiguous():
        raise ValueError("dwconv3x3 needs fp16 tensors and the contiguous [9, C] matrix of dwconv3x3_tap_major()")
    sx = _slot_stride(x, b, c, hh, ww, "x")
    if out is None:
        out = torch.empty((b, c, hh, ww), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    so = _slot_stride(out, b, c, hh, ww, "out")
    h = _handle_for(x)
    check(h.lib.fsd_dwconv3x3(h.h, x.data_ptr(), sx, b, hh, ww, weight_taps.data_ptr(), bias.data_ptr(), out.data_ptr(), so, c,
                              _ACT[act], float(slope), _TORCH_DTYPE[x.dtype], _stream_ptr(x.device)), "fsd_dwconv3x3")
    return out


def sppf_pool_(buf: torch.Tensor) -> torch.Tensor:
    """(a5) fills channel slots 1..3 of the dense channels-last [N,4c,H,W] fp16 buffer with the cascaded 5x5 max pools of
    slot 0 (ultralytics SPPF), one launch."""
    _require_cuda(buf, "buf")
    n, c4, hh, ww = buf.shape
    if buf.dtype != torch.float16 or not buf.is_contiguous(memory_format=torch.channels_last) or c4 % 32:
        raise ValueError("sppf_pool_ needs a dense channels_last fp16 [N,4c,H,W] buffer with c % 8 == 0")
    h = _handle_for(buf)
    check(h.lib.fsd_sppf_pool(h.h, buf.data_ptr(), n, hh, ww, c4 // 4, _TORCH_DTYPE[buf.dtype], _stream_ptr(buf.device)),
          "fsd_sppf_pool")
    return buf


def upsample2x_concat(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """(a5) cat(interpolate(a, 2x nearest), b, dim=1) for channels-last CUDA tensors, one pass."""
    _require_cuda(a, "a")
    n, ca, ah, aw = a.shape
    nb, cb, bh, bw = b.shape
    if (nb, bh, bw) != (n, 2 * ah, 2 * aw) or a.dtype != b.dtype:
        raise ValueError("upsample2x_concat: b must be [N, Cb, 2h, 2w] of a's dtype")
    if not (a.is_contiguous(memory_format=torch.channels_last) and b.is_contiguous(memory_format=torch.channels_last)):
        raise ValueError("upsample2x_concat needs dense channels_last tensors")
    out = torch.empty((n, ca + cb, bh, bw), dtype=a.dtype, device=a.device, memory_format=torch.channels_last)
    h = _handle_for(a)
    check(h.lib.fsd_upsample2x_concat(h.h, a.data_ptr(), b.data_ptr(), out.data_ptr(), n, ah, aw, ca, cb,
                                      _TORCH_DTYPE[a.dtype], _stream_ptr(a.device)), "fsd_upsample2x_concat")
    return out


# ---- Kernel 4 -------------------------------------------------------------------------------------------
def esrgan_tile_table(H: int, W: int, scale: int, tile: int, tile_pad: int = 10, pre_pad: int = 0):
    """Host tile table [T,12] int32 + (padded_h, padded_w)."""
    lib = _cabi.load_library()
    n = C.c_int(0)
    hw = (C.c_int32 * 2)()
    check(lib.fsd_esrgan_tile_table(H, W, scale, tile, tile_pad, pre_pad, None, 0, C.byref(n), hw), "fsd_esrgan_tile_table")
    tab = np.zeros((max(1, n.value), 12), dtype=np.int32)
    check(lib.fsd_esrgan_tile_table(H, W, scale, tile, tile_pad, pre_pad, tab.ctypes.data_as(_cabi.c_i32p),
                                    n.value, C.byref(n), hw), "fsd_esrgan_tile_table")
    return tab[: n.value], (int(hw[0]), int(hw[1]))


def _off64(row, k):
    return int(np.uint32(row[k])) | (int(row[k + 1]) << 32)


def esrgan_tile_elems(table: np.ndarray, scale: int = 1, out: bool = False) -> int:
    """Elements of one image's packed tile buffer (input tiles, or network outputs when out=True)."""
    last = table[-1]
    n = 3 * int(last[2]) * int(last[3]) * (scale * scale if out else 1)
    return _off64(last, 10 if out else 8) + (n + 7) // 8 * 8


def esrgan_crop(img: torch.Tensor, table: np.ndarray, scale: int, pre_pad: int = 0, dtype=torch.float16,
                tab_dev: torch.Tensor | None = None, tiles: torch.Tensor | None = None):
    """Kernel 4a: img [H,W,3] — or a batch [N,H,W,3] — uint8 BGR (CUDA, rows contiguous) -> (packed tile buffer
    [elems] or [N, elems], table on device).  One launch crops the whole batch with the same tile table.
    Pass `tab_dev` / `tiles` from a previous call to reuse the uploaded table and the tile buffer."""
    _require_cuda(img, "image")
    batched = img.dim() == 4
    im = img if batched else img[None]
    N, H, W = int(im.shape[0]), int(im.shape[1]), int(im.shape[2])
    assert im.dtype == torch.uint8 and im.stride(3) == 1 and im.stride(2) == 3
    total = esrgan_tile_elems(table)
    if tiles is None:
        tiles = torch.empty((N, total) if batched else (total,), dtype=dtype, device=img.device)
    tl = tiles if tiles.dim() == 2 else tiles[None]
    assert tl.shape[0] == N and tl.stride(1) == 1 and tl.shape[1] >= total
    tab_host = np.ascontiguousarray(table, dtype=np.int32)
    if tab_dev is None:
        tab_dev = torch.from_numpy(tab_host).to(img.device)
    h = _handle_for(img)
    check(h.lib.fsd_esrgan_crop(h.h, im.data_ptr(), H, W, im.stride(1), H + pre_pad, W + pre_pad,
                                tab_dev.data_ptr(), tab_host.ctypes.data, len(tab_host), _TORCH_DTYPE[tl.dtype],
                                tl.data_ptr(), N, im.stride(0) if N > 1 else 0, tl.stride(0) if N > 1 else 0,
                                _stream_ptr(img.device)), "fsd_esrgan_crop")
    return tiles, tab_dev


def tile_view(buf: torch.Tensor, row, scale: int = 1, out: bool = False) -> torch.Tensor:
    """[1,3,h,w] view of one tile inside a packed buffer (input tiles: out=False; network outputs: out=True)."""
    off = _off64(row, 10 if out else 8)
    h, w = int(row[3]) * scale, int(row[2]) * scale
    return buf[off: off + 3 * h * w].view(1, 3, h, w)


def esrgan_out_buffer(table: np.ndarray, scale: int, dtype, device, n_images: int | None = None) -> torch.Tensor:
    total = esrgan_tile_elems(table, scale, out=True)
    return torch.empty((total,) if n_images is None else (n_images, total), dtype=dtype, device=device)


def esrgan_stitch(tiles_out: torch.Tensor, table: np.ndarray, tab_dev: torch.Tensor, scale: int, H: int, W: int,
                  out: torch.Tensor | None = None) -> torch.Tensor:
    """Kernel 4b: packed network outputs [elems] -> [H*scale, W*scale, 3] uint8 BGR, or a batch [N, elems] ->
    [N, H*scale, W*scale, 3] in one launch."""
    _require_cuda(tiles_out, "tile outputs")
    oh, ow = H * scale, W * scale
    batched = tiles_out.dim() == 2
    tl = tiles_out if batched else tiles_out[None]
    N = int(tl.shape[0])
    if out is None:
        out = torch.empty((N, oh, ow, 3) if batched else (oh, ow, 3), dtype=torch.uint8, device=tiles_out.device)
    o = out if out.dim() == 4 else out[None]
    assert o.shape[0] == N and o.stride(3) == 1 and o.stride(2) == 3 and tl.stride(1) == 1
    tab_host = np.ascontiguousarray(table, dtype=np.int32)
    h = _handle_for(tiles_out)
    check(h.lib.fsd_esrgan_stitch(h.h, tl.data_ptr(), tab_dev.data_ptr(), tab_host.ctypes.data, len(tab_host),
                                  scale, _TORCH_DTYPE[tl.dtype], o.data_ptr(), oh, ow, o.stride(1), N,
                                  tl.stride(0) if N > 1 else 0, o.stride(0) if N > 1 else 0,
                                  _stream_ptr(tiles_out.device)), "fsd_esrgan_stitch")
    return out


# ---- evaluation helpers (f1, f2) -------------------------------------------------------------------------
def bbox_overlaps_p1(boxes: torch.Tensor, query: torch.Tensor) -> torch.Tensor:
    """WIDER-FACE bbox_overlaps ('+1' convention): boxes [N,4], query [K,4] float64 CUDA -> [N,K] float64."""
    _require_cuda(boxes, "boxes")
    boxes = boxes.to(torch.float64).contiguous()
    query = query.to(device=boxes.device, dtype=torch.float64).contiguous()
    out = torch.zeros((boxes.shape[0], query.shape[0]), dtype=torch.float64, device=boxes.device)
    h = _handle_for(boxes)
    check(h.lib.fsd_bbox_overlaps_p1(h.h, boxes.data_ptr(), int(boxes.shape[0]), query.data_ptr(),
                                     int(query.shape[0]), out.data_ptr(), _stream_ptr(boxes.device)),
          "fsd_bbox_overlaps_p1")
    return out


def widerface_pr_curve(pred: torch.Tensor, pred_off: torch.Tensor, gt: torch.Tensor, gt_off: torch.Tensor,
                       evaluate: torch.Tensor, thresh: torch.Tensor, iou_thresh: float = 0.5):
    """(f1) one launch: greedy matching of every image + PR accumulation.  pred [Np,5] / gt [Ng,4] float64 xywh(+score)
    rows with int32 CSR offsets [G+1], evaluate [Ng] int32, thresh [T] float64 (all CUDA).
    Returns (pr_curve [T,2], pred_recall [Np], proposal [Np]) float64."""
    _require_cuda(pred_off, "offsets")
    dev = pred_off.device
    G, n_pred, n_gt, T = int(pred_off.shape[0]) - 1, int(pred.shape[0]), int(gt.shape[0]), int(thresh.shape[0])
    for t, dt in ((pred, torch.float64), (gt, torch.float64), (thresh, torch.float64), (pred_off, torch.int32),
                  (gt_off, torch.int32), (evaluate, torch.int32)):
        assert t.dtype == dt and t.is_contiguous() and t.device == dev
    pr_curve = torch.empty((T, 2), dtype=torch.float64, device=dev)
    pred_recall = torch.zeros((n_pred,), dtype=torch.float64, device=dev)
    proposal = torch.ones((n_pred,), dtype=torch.float64, device=dev)
    h = get_handle(dev.index if dev.index is not None else torch.cuda.current_device())
    nbytes = int(h.lib.fsd_widerface_scratch_bytes(n_pred, n_gt))
    scratch = torch.empty((nbytes // 8 + 1,), dtype=torch.float64, device=dev)
    check(h.lib.fsd_widerface_pr_curve(h.h, _ptr(pred), pred_off.data_ptr(), _ptr(gt), gt_off.data_ptr(), _ptr(evaluate), G,
                                       n_pred, n_gt, float(iou_thresh), thresh.data_ptr(), T, pred_recall.data_ptr(),
                                       proposal.data_ptr(), scratch.data_ptr(), nbytes, pr_curve.data_ptr(), _stream_ptr(dev)),
          "fsd_widerface_pr_curve")
    return pr_curve, pred_recall, proposal


def attach_keypoints(merged: torch.Tensor, m_off, m_cnt, dets: torch.Tensor, d_off, d_cnt) -> torch.Tensor:
    """(f2) per merged box -> global detection row whose key-points the reference would attach (-1: none)."""
    _require_cuda(merged, "merged boxes")
    src = torch.full((merged.shape[0],), -1, dtype=torch.int32, device=merged.device)
    h = _handle_for(merged)
    check(h.lib.fsd_attach_keypoints(h.h, merged.data_ptr(), merged.stride(0), m_off.data_ptr(), m_cnt.data_ptr(),
                                     dets.data_ptr(), dets.stride(0), d_off.data_ptr(), d_cnt.data_ptr(),
                                     int(m_off.shape[0]), src.data_ptr(), _stream_ptr(merged.device)),
          "fsd_attach_keypoints")
    return src


def jpeg_info(data: bytes):
    """(width, height, channels) of a JPEG stream, parsed by nvJPEG on the host."""
    lib = _cabi.load_library()
    w, hh, c = C.c_int(0), C.c_int(0), C.c_int(0)
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    check(lib.fsd_jpeg_info(buf, len(data), C.byref(w), C.byref(hh), C.byref(c)), "fsd_jpeg_info")
    return w.value, hh.value, c.value
