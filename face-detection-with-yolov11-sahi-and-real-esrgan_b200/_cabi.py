"""ctypes binding of include/fsd_b200.h.

This is the only place the product touches native code.  There is NO CPU fallback: if the library cannot
be loaded, or no sm_100 device is present when a handle is requested, the call raises.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

from . import _build

c_i32p = C.POINTER(C.c_int32)
c_f32p = C.POINTER(C.c_float)
vp = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/fsd_b200.h declares (tests check this).
SIGNATURES = {
    "fsd_version": (C.c_char_p, []),
    "fsd_last_error": (C.c_char_p, []),
    "fsd_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "fsd_destroy": (C.c_int, [vp]),
    "fsd_launch_count": (C.c_int64, [vp]),
    "fsd_slice_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, c_i32p, C.c_int,
                                 C.POINTER(C.c_int)]),
    "fsd_letterbox_geometry": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, c_i32p, C.POINTER(C.c_double)]),
    "fsd_gather_letterbox": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, vp, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "fsd_pose_decode": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), c_i32p, C.c_int, C.c_int,
                                  C.c_int, C.c_float, vp, C.c_int, vp, vp]),
    "fsd_merge_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int, C.c_int]),
    "fsd_merge": (C.c_int, [vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp, C.c_int, C.c_int,
                            C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp,
                            vp, vp, vp, vp, C.c_int64, vp]),
    "fsd_finalize_dets": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, vp,
                                    C.c_int, vp, vp]),
    "fsd_pack_results": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp]),
    "fsd_bias_act_inplace": (C.c_int, [vp, vp, vp, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int, vp]),
    "fsd_kernel_timing_enable": (C.c_int, [vp, C.c_uint]),
    "fsd_kernel_timing_read": (C.c_int, [vp, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int)]),
    "fsd_bias_act": (C.c_int, [vp, vp, vp, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, C.c_int, vp, C.c_int64, C.c_int, C.c_int,
                               C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int, vp]),
    "fsd_stem_conv": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]),
    "fsd_pointwise_conv_supported": (C.c_int, [C.c_int, C.c_int]),
    "fsd_pointwise_conv": (C.c_int, [vp, vp, C.c_int64, vp, vp, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, C.c_int,
                                     C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, vp]),
    "fsd_conv3x3_supported": (C.c_int, [C.c_int, C.c_int]),
    "fsd_conv3x3": (C.c_int, [vp, vp, C.c_int64, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_int64, vp, C.c_int64, C.c_int, C.c_int,
                              C.c_int, C.c_float, C.c_int, vp]),
    "fsd_conv2x2_supported": (C.c_int, [C.c_int, C.c_int]),
    "fsd_conv2x2": (C.c_int, [vp, vp, C.c_int64, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, vp]),
    "fsd_dwconv3x3": (C.c_int, [vp, vp, C.c_int64, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int, vp]),
    "fsd_sppf_pool": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "fsd_upsample2x_concat": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "fsd_esrgan_tile_table": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_i32p, C.c_int,
                                        C.POINTER(C.c_int), c_i32p]),
    "fsd_esrgan_crop": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int,
                                  vp, C.c_int, C.c_int64, C.c_int64, vp]),
    "fsd_esrgan_stitch": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int64,
                                    C.c_int, C.c_int64, C.c_int64, vp]),
    "fsd_bbox_overlaps_p1": (C.c_int, [vp, vp, C.c_int, vp, C.c_int, vp, vp]),
    "fsd_widerface_scratch_bytes": (C.c_int64, [C.c_int64, C.c_int64]),
    "fsd_widerface_pr_curve": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int64, C.c_int64, C.c_double, vp, C.c_int, vp, vp,
                                         vp, C.c_int64, vp, vp]),
    "fsd_jpeg_info": (C.c_int, [vp, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "fsd_jpeg_decode": (C.c_int, [vp, vp, C.c_int64, C.c_int, vp, C.c_int64, C.c_int, C.c_int, vp]),
    "fsd_attach_keypoints": (C.c_int, [vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, vp]),
}

FSD_F16, FSD_F32 = 0, 1
FSD_NMS, FSD_GREEDYNMM, FSD_NMM = 0, 1, 2
FSD_IOU, FSD_IOS = 0, 1
FSD_KERNEL_GATHER, FSD_KERNEL_DECODE, FSD_KERNEL_MERGE, FSD_KERNEL_ESRGAN_CROP, FSD_KERNEL_BIAS_ACT = 1, 2, 3, 4, 5
FSD_KERNEL_STEM, FSD_KERNEL_POINTWISE, FSD_KERNEL_FINALIZE, FSD_KERNEL_ATTACH, FSD_KERNEL_PACK = 6, 7, 8, 9, 10
FSD_KERNEL_ESRGAN_STITCH, FSD_KERNEL_SPPF, FSD_KERNEL_CONV3X3, FSD_KERNEL_DWCONV = 11, 12, 13, 14
KERNEL_NAMES = {1: "k1_gather", 2: "k2a_decode", 3: "k3_merge", 4: "k4_crop", 5: "k5_bias_act", 6: "k6_stem", 7: "k7_pointwise",
                8: "k2b_finalize", 9: "attach_keypoints", 10: "pack", 11: "k4_stitch", 12: "k5_sppf", 13: "k10_conv3x3", 14: "k11_dwconv3x3"}
FSD_PLANAR, FSD_CHANNELS_LAST = 0, 1

_lib = None
MISSING: list[str] = []
_lib_lock = threading.Lock()


class FsdError(RuntimeError):
    pass


def library_path() -> Path:
    return _build.LIB_PATH


def load_library(build_if_missing: bool = True) -> C.CDLL:
    """dlopen libfsd_b200.so (building it in-tree with nvcc when stale) and attach prototypes."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if build_if_missing:
            _build.build()
        path = library_path()
        if not path.exists():
            raise FsdError(f"{path} is missing: build it with `python __graft_entry__.py` (no CPU fallback exists)")
        lib = C.CDLL(str(path))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name, None)
            if fn is None:  # header/implementation drift; tests assert MISSING stays empty
                MISSING.append(name)
                continue
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load_library().fsd_last_error().decode("utf-8", "replace")
        raise FsdError(f"{what} failed with status {rc}: {msg}")


class Handle:
    """Owns one fsd_handle_t bound to a CUDA device."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = vp()
        check(self.lib.fsd_create(int(device), C.byref(h)), "fsd_create")
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.fsd_destroy(self.h)
            self.h = None

    def __del__(self):  # best effort
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.fsd_launch_count(self.h))

    def timing_enable(self, kernels=(FSD_KERNEL_GATHER,)) -> None:
        """Start (clearing earlier samples) the in-library CUDA-event timing of the given FSD_KERNEL_* ids; () stops it."""
        mask = 0
        for k in kernels or ():
            mask |= 1 << int(k)
        check(self.lib.fsd_kernel_timing_enable(self.h, mask), "fsd_kernel_timing_enable")

    def timing_read(self):
        """[(kernel id, units, tag, ms)] for every instrumented launch since timing_enable(True); waits for them."""
        n = C.c_int(0)
        check(self.lib.fsd_kernel_timing_read(self.h, None, 0, C.byref(n)), "fsd_kernel_timing_read")
        if n.value == 0:
            return []
        buf = (C.c_double * (4 * n.value))()
        check(self.lib.fsd_kernel_timing_read(self.h, buf, n.value, C.byref(n)), "fsd_kernel_timing_read")
        return [(int(buf[4 * i]), int(buf[4 * i + 1]), int(buf[4 * i + 2]), float(buf[4 * i + 3])) for i in range(n.value)]


_handles: dict[int, Handle] = {}


def get_handle(device: int = 0) -> Handle:
    """Process-wide handle per device index."""
    h = _handles.get(device)
    if h is None:
        h = Handle(device)
        _handles[device] = h
    return h


# ---- host-only planners (usable without a GPU) -----------------------------------------------------

def slice_plan(image_h: int, image_w: int, slice_h: int, slice_w: int, overlap_h: float, overlap_w: float):
    """[[x0,y0,x1,y1], ...] — sahi get_slice_bboxes semantics, computed by the C library."""
    lib = load_library()
    n = C.c_int(0)
    check(lib.fsd_slice_plan(image_h, image_w, slice_h, slice_w, overlap_h, overlap_w, None, 0, C.byref(n)),
          "fsd_slice_plan")
    buf = (C.c_int32 * (4 * max(1, n.value)))()
    check(lib.fsd_slice_plan(image_h, image_w, slice_h, slice_w, overlap_h, overlap_w, buf, n.value, C.byref(n)),
          "fsd_slice_plan")
    return [[buf[4 * i + k] for k in range(4)] for i in range(n.value)]


def letterbox_geometry(src_h: int, src_w: int, imgsz: int = 1024, stride: int = 32) -> dict:
    lib = load_library()
    g = (C.c_int32 * 8)()
    gain = C.c_double(0)
    check(lib.fsd_letterbox_geometry(src_h, src_w, imgsz, stride, g, C.byref(gain)), "fsd_letterbox_geometry")
    return dict(new_w=g[0], new_h=g[1], left=g[2], top=g[3], out_w=g[4], out_h=g[5], mode=g[6], gain=gain.value)
