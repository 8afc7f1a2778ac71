"""Real-ESRGAN tiled enhancement with the reference's wrapper contracts.

`RealESRGANer` mirrors [EXT realesrgan 0.3.0] (constructor, `.enhance(img, outscale) -> (img, mode)`, the `tile_size`
attribute) and `FaceEnhancer` mirrors utils/enhancer.py:21-302 (`enhance_image(img) -> (img, ok)` never raises).
Tile crop (+ reflect pre-/mod-pad, /255, BGR->RGB, fp16) and stitch (interior copy, clamp, *255, round-half-even,
RGB->BGR, uint8) are Kernel 4; RRDBNet stays PyTorch and runs batched over same-shaped tiles instead of tile by tile."""
from __future__ import annotations

import os
from collections import defaultdict

import numpy as np
import torch

from . import _cabi, ops
from .backbones.rrdbnet import RRDBNet


class RealESRGANer:
    def __init__(self, scale, model_path=None, dni_weight=None, model=None, tile=0, tile_pad=10, pre_pad=10,
                 half=False, device=None, gpu_id=None, max_tile_batch: int = 8):
        if not torch.cuda.is_available():
            raise _cabi.FsdError("fsd_b200.RealESRGANer runs on CUDA only (no CPU fallback)")
        self.scale, self.tile_size, self.tile_pad, self.pre_pad, self.half = scale, tile, tile_pad, pre_pad, half
        self.mod_scale = None
        self.device = torch.device(device) if device is not None else torch.device(f"cuda:{gpu_id or 0}")
        if model is None:
            model = RRDBNet(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=23, num_grow_ch=32, scale=scale)
        self.weights = "random-init"
        if model_path:  # like RealESRGANer: a given path is loaded ('params_ema' preferred) or the call raises
            from .checkpoints import load_rrdbnet_state

            model.load_state_dict(load_rrdbnet_state(str(model_path)), strict=True)
            self.weights = str(model_path)
        model = model.to(self.device).eval()
        for p in model.parameters():
            p.requires_grad_(False)
        self.model = model.half() if half else model.float()
        self.max_tile_batch = max_tile_batch
        self._tab_cache = {}

    @torch.no_grad()
    def enhance_device(self, img_dev: torch.Tensor) -> torch.Tensor:
        """[H,W,3] uint8 BGR on the device -> [H*s,W*s,3] uint8 BGR on the device (no host round trip).
        A batch [N,H,W,3] of same-sized frames is cropped and stitched by ONE launch each and its same-shaped tiles share
        RRDBNet batches across frames (the reference enhances frame by frame, tile by tile)."""
        batched = img_dev.dim() == 4
        frames = img_dev if batched else img_dev[None]
        N, H, W = int(frames.shape[0]), int(frames.shape[1]), int(frames.shape[2])
        s = self.scale
        tile = self.tile_size if self.tile_size > 0 else max(H, W) + self.pre_pad + 4  # tile 0: one tile, no halo
        table, _ = ops.esrgan_tile_table(H, W, s, tile, self.tile_pad if self.tile_size > 0 else 0, self.pre_pad)
        dtype = torch.float16 if self.half else torch.float32
        key = (H, W)
        cached = self._tab_cache.get(key)
        tiles, tab_dev = ops.esrgan_crop(frames, table, s, self.pre_pad, dtype, tab_dev=cached)
        self._tab_cache[key] = tab_dev
        outbuf = ops.esrgan_out_buffer(table, s, dtype, img_dev.device, n_images=N)
        groups = defaultdict(list)
        for n in range(N):
            for i, row in enumerate(table):
                groups[(int(row[3]), int(row[2]))].append((n, i))
        for (_, _), idxs in groups.items():  # same-shaped tiles go through RRDBNet as one batch
            for a in range(0, len(idxs), self.max_tile_batch):
                part = idxs[a:a + self.max_tile_batch]
                views = [ops.tile_view(tiles[n], table[i]) for n, i in part]
                x = torch.cat(views, 0) if len(part) > 1 else views[0]
                with ops.cudnn_benchmark():  # tile shapes repeat frame after frame
                    y = self.model(x)
                for j, (n, i) in enumerate(part):
                    ops.tile_view(outbuf[n], table[i], s, out=True).copy_(y[j:j + 1])
        out = ops.esrgan_stitch(outbuf, table, tab_dev, s, H, W)
        return out if batched else out[0]

    def enhance(self, img, outscale=None, alpha_upsampler="realesrgan"):
        if not isinstance(img, np.ndarray) or img.ndim != 3 or img.shape[2] != 3 or img.dtype != np.uint8:
            raise TypeError("fsd_b200.RealESRGANer.enhance takes an HWC uint8 BGR image (the only form the reference passes)")
        if outscale is not None and float(outscale) != float(self.scale):
            raise NotImplementedError("outscale != scale (Lanczos post-resize) is never used by the reference")
        dev_img = torch.from_numpy(np.ascontiguousarray(img)).to(self.device)
        out = self.enhance_device(dev_img)
        return out.cpu().numpy(), "RGB"


class FaceEnhancer:
    def __init__(self, model_name="RealESRGAN_x4plus", model_path=None, scale=4, tile=400, half=True,
                 allow_random_init=False):
        """As utils/enhancer.py:21-60: the constructor raises when the weights cannot be found — unless
        `allow_random_init=True` (tests / bench: there is no network for the real weights)."""
        self.model_name, self.scale, self.tile, self.half = model_name, scale, tile, half
        self.allow_random_init = allow_random_init
        self.upsampler = None
        self.device = self._check_device()
        self._setup_model(model_name, model_path)

    def _check_device(self):
        return "cuda" if torch.cuda.is_available() else "cpu"

    def _find_model_path(self, model_name):
        for cand in (f"models/{model_name}.pth", f"weights/{model_name}.pth", f"{model_name}.pth"):
            if os.path.exists(cand):
                return os.path.abspath(cand)
        return None

    def _setup_model(self, model_name, model_path):
        if self.device == "cpu":
            raise _cabi.FsdError("fsd_b200.FaceEnhancer needs a CUDA device: the tile crop/stitch kernels have no CPU fallback")
        if model_path is None:
            model_path = self._find_model_path(model_name)
        if model_path is None or not os.path.isfile(str(model_path)):
            # the reference cannot continue either: RealESRGANer(model_path=None) raises inside the constructor
            # (utils/enhancer.py:131-186 retries once, then re-raises); there is no network here to download weights
            if not self.allow_random_init:
                raise FileNotFoundError(f"Real-ESRGAN weights for {model_name} not found ({model_path}); pass model_path=... "
                                        "(or allow_random_init=True for tests / benchmarks)")
            model_path = None
        blocks = 6 if "anime_6B" in model_name else 23
        if "x2" in model_name and "anime_6B" not in model_name:
            self.scale = 2
        torch.manual_seed(0)  # deterministic random-init when no weight file is present (no network here)
        model = RRDBNet(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=blocks, num_grow_ch=32, scale=self.scale)
        self.upsampler = RealESRGANer(scale=self.scale, model_path=model_path, dni_weight=None, model=model,
                                      tile=self.tile, tile_pad=10, pre_pad=0, half=self.half, gpu_id=0)

    def enhance_image(self, image):
        if self.upsampler is None:
            return image, False
        try:
            from PIL import Image

            if isinstance(image, Image.Image):
                image = np.ascontiguousarray(np.array(image)[:, :, ::-1])  # RGB -> BGR
            if image is None or image.size == 0:
                return image, False
            h, w = image.shape[:2]
            if h < 4 or w < 4:
                return image, False
            out, _ = self.upsampler.enhance(image, outscale=self.scale)
            return out, True
        except Exception as e:  # the reference swallows everything and reports failure through the flag
            print(f" Enhancement failed: {type(e).__name__}: {e}")
            return image, False

    def enhance_face_crop(self, crop_path, output_path, quality=95):
        import cv2

        info = dict(original_path=crop_path, output_path=output_path, original_size=None, enhanced_size=None,
                    scale_factor=self.scale, success=False)
        if not os.path.exists(crop_path):
            return False, info
        img = cv2.imread(crop_path, cv2.IMREAD_COLOR)
        if img is None:
            return False, info
        info["original_size"] = (img.shape[1], img.shape[0])
        out, ok = self.enhance_image(img)
        if not ok:
            return False, info
        info["enhanced_size"] = (out.shape[1], out.shape[0])
        os.makedirs(os.path.dirname(output_path) or ".", exist_ok=True)
        ext = os.path.splitext(output_path)[1].lower()
        args = [cv2.IMWRITE_JPEG_QUALITY, quality] if ext in (".jpg", ".jpeg") else []
        if not cv2.imwrite(output_path, out, args):
            return False, info
        info["success"] = True
        return True, info

    def get_model_info(self):
        return dict(model_name=self.model_name, scale=self.scale, tile=self.tile, half_precision=self.half,
                    device=self.device, is_loaded=self.upsampler is not None, torch_version=torch.__version__,
                    cuda_available=torch.cuda.is_available())
