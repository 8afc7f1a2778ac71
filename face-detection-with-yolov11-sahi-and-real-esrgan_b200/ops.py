"""Torch-tensor front ends of the C-ABI kernels.

PyTorch is used here for device memory, streams and dtype bookkeeping only; every byte of arithmetic on the hot
path is done by the hand-written sm_100a kernels in csrc/.  All functions raise if CUDA is unavailable.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi
from ._cabi import FSD_F16, FSD_F32, check, get_handle

_TORCH_DTYPE = {torch.float16: FSD_F16, torch.float32: FSD_F32}


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

    def view(self, index: int) -> torch.Tensor:
        """[H, W, 3] view of slot `index` (strided when the pitch is padded)."""
        return self.buf[index, :, : self.w * 3].unflatten(1, (self.w, 3))

    @classmethod
    def from_numpy(cls, images, device="cuda") -> "ImagePool":
        h, w = images[0].shape[:2]
        pool = cls(len(images), h, w, device)
        for i, im in enumerate(images):
            pool.upload(i, im, non_blocking=False)
        return pool


def gather_letterbox(pool: ImagePool, entries: torch.Tensor, src_w: int, src_h: int, imgsz: int = 1024,
                     stride: int = 32, reverse_channels: bool = True, dtype=torch.float16,
                     out: torch.Tensor | None = None) -> torch.Tensor:
    """Kernel 1: entries int32 [B,3] (image_index, x0, y0) -> [B,3,out_h,out_w] network input."""
    _require_cuda(pool.buf, "image pool")
    g = _cabi.letterbox_geometry(src_h, src_w, imgsz, stride)
    B = int(entries.shape[0])
    if out is None:
        out = torch.empty((B, 3, g["out_h"], g["out_w"]), dtype=dtype, device=pool.device)
    if entries.dtype != torch.int32 or not entries.is_cuda or not entries.is_contiguous():
        entries = entries.to(device=pool.device, dtype=torch.int32).contiguous()
    h = _handle_for(pool.buf)
    check(h.lib.fsd_gather_letterbox(h.h, pool.buf.data_ptr(), pool.n, pool.h, pool.w, pool.pitch,
                                     pool.image_pitch, entries.data_ptr(), B, src_w, src_h, imgsz, stride,
                                     1 if reverse_channels else 0, _TORCH_DTYPE[out.dtype], out.data_ptr(),
                                     _stream_ptr(pool.device)), "fsd_gather_letterbox")
    return out
