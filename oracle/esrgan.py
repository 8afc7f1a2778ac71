"""Oracle for Kernel 4: realesrgan.RealESRGANer tile crop / stitch (test infrastructure, see oracle/__init__).

Restates [EXT realesrgan==0.3.0] `RealESRGANer.{enhance,pre_process,tile_process,post_process}` as recorded in
SURVEY.md App. A.5 — reached from the reference at utils/enhancer.py:138-156 (construction) and :214 (enhance).
Parity unpinned (upstream package absent).  The up-sampling model is any callable tensor -> tensor (RRDBNet in the
reference; an exact nearest-neighbour "identity up-sampler" in the known-answer tests).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


class RealESRGANer:
    def __init__(self, scale, model_path=None, dni_weight=None, model=None, tile=0, tile_pad=10, pre_pad=10,
                 half=False, device=None, gpu_id=None):
        self.scale = scale
        self.tile_size = tile  # NB: stored as tile_size; the reference's OOM retry assigns `.tile` (a no-op)
        self.tile_pad = tile_pad
        self.pre_pad = pre_pad
        self.mod_scale = None
        self.half = half
        self.device = torch.device(device if device is not None else "cpu")
        self.model = model.to(self.device).eval() if hasattr(model, "to") else model
        if self.half and hasattr(self.model, "half"):
            self.model = self.model.half()

    def pre_process(self, img):
        img = torch.from_numpy(np.transpose(img, (2, 0, 1))).float()
        self.img = img.unsqueeze(0).to(self.device)
        if self.half:
            self.img = self.img.half()
        if self.pre_pad != 0:
            self.img = F.pad(self.img, (0, self.pre_pad, 0, self.pre_pad), "reflect")
        if self.scale == 2:
            self.mod_scale = 2
        elif self.scale == 1:
            self.mod_scale = 4
        if self.mod_scale is not None:
            self.mod_pad_h, self.mod_pad_w = 0, 0
            _, _, h, w = self.img.size()
            if h % self.mod_scale != 0:
                self.mod_pad_h = self.mod_scale - h % self.mod_scale
            if w % self.mod_scale != 0:
                self.mod_pad_w = self.mod_scale - w % self.mod_scale
            self.img = F.pad(self.img, (0, self.mod_pad_w, 0, self.mod_pad_h), "reflect")

    def process(self):
        self.output = self.model(self.img)

    def tile_table(self):
        """[(px0,py0,px1,py1, ix0,iy0,ix1,iy1)] in tile_process order (row-major)."""
        _, _, height, width = self.img.shape
        rows = []
        for y in range(math.ceil(height / self.tile_size)):
            for x in range(math.ceil(width / self.tile_size)):
                ix0, iy0 = x * self.tile_size, y * self.tile_size
                ix1, iy1 = min(ix0 + self.tile_size, width), min(iy0 + self.tile_size, height)
                px0, px1 = max(ix0 - self.tile_pad, 0), min(ix1 + self.tile_pad, width)
                py0, py1 = max(iy0 - self.tile_pad, 0), min(iy1 + self.tile_pad, height)
                rows.append((px0, py0, px1, py1, ix0, iy0, ix1, iy1))
        return rows

    def tile_process(self):
        batch, channel, height, width = self.img.shape
        s = self.scale
        self.output = self.img.new_zeros((batch, channel, height * s, width * s))
        self.last_tiles = []
        for (px0, py0, px1, py1, ix0, iy0, ix1, iy1) in self.tile_table():
            in_tile = self.img[:, :, py0:py1, px0:px1]
            with torch.no_grad():
                out_tile = self.model(in_tile)
            self.last_tiles.append((in_tile, out_tile))
            tx0, ty0 = (ix0 - px0) * s, (iy0 - py0) * s
            tx1, ty1 = tx0 + (ix1 - ix0) * s, ty0 + (iy1 - iy0) * s
            self.output[:, :, iy0 * s:iy1 * s, ix0 * s:ix1 * s] = out_tile[:, :, ty0:ty1, tx0:tx1]

    def post_process(self):
        if self.mod_scale is not None:
            _, _, h, w = self.output.size()
            self.output = self.output[:, :, 0:h - self.mod_pad_h * self.scale, 0:w - self.mod_pad_w * self.scale]
        if self.pre_pad != 0:
            _, _, h, w = self.output.size()
            self.output = self.output[:, :, 0:h - self.pre_pad * self.scale, 0:w - self.pre_pad * self.scale]
        return self.output

    @torch.no_grad()
    def enhance(self, img, outscale=None, alpha_upsampler="realesrgan"):
        """img: HWC uint8 BGR (the only form the reference passes) -> (HWC uint8 BGR, 'RGB')."""
        h_in, w_in = img.shape[0:2]
        img = img.astype(np.float32)
        max_range = 65535 if np.max(img) > 256 else 255
        img = img / max_range
        assert img.ndim == 3 and img.shape[2] == 3, "gray / alpha inputs are outside the reference's usage"
        img = img[:, :, ::-1].copy()  # cv2.COLOR_BGR2RGB on float32
        self.pre_process(img)
        if self.tile_size > 0:
            self.tile_process()
        else:
            self.process()
        out = self.post_process()
        out = out.data.squeeze().float().cpu().clamp_(0, 1).numpy()
        out = np.transpose(out[[2, 1, 0], :, :], (1, 2, 0))
        out = (out * 255.0).round().astype(np.uint8)
        if outscale is not None and outscale != float(self.scale):
            import cv2

            out = cv2.resize(out, (int(w_in * outscale), int(h_in * outscale)), interpolation=cv2.INTER_LANCZOS4)
        return out, "RGB"


class NearestUpsampler(torch.nn.Module):
    """Exact 'identity' up-sampler for known-answer tests: nearest-neighbour x scale (x2 nets see pixel-unshuffled input
    upstream, but as a black box the model maps [1,3,h,w] -> [1,3,h*s,w*s])."""

    def __init__(self, scale):
        super().__init__()
        self.scale = scale

    def forward(self, x):
        return F.interpolate(x, scale_factor=self.scale, mode="nearest")
