"""Deterministic synthetic inputs of the BASELINE.json shapes: WIDER-FACE-shaped images with pasted
elliptical "faces" and their ground-truth boxes.  There is no network for datasets, so every bench/test input comes
from here; image i is a pure function of (seed + i)."""
from __future__ import annotations

import numpy as np


def make_image(index: int, height: int = 768, width: int = 1024, seed: int = 1234, mean_faces: float = 12.0,
               face_px=(6, 200)):
    """-> (HWC uint8 image, gt [K,4] float xywh).  Low-pass noise background + K~Poisson(mean_faces) ellipses."""
    rng = np.random.default_rng(seed + index)
    bg = rng.integers(0, 256, (height + 2, width + 2, 3), dtype=np.uint16)
    acc = np.zeros((height, width, 3), dtype=np.uint16)
    for dy in range(3):  # 3x3 box filter
        for dx in range(3):
            acc += bg[dy:dy + height, dx:dx + width]
    img = (acc // 9).astype(np.uint8)
    k = int(rng.poisson(mean_faces))
    gts = []
    yy, xx = np.mgrid[0:height, 0:width]
    for _ in range(k):
        w = float(np.exp(rng.uniform(np.log(face_px[0]), np.log(face_px[1]))))
        h = w * rng.uniform(1.1, 1.4)
        if h >= height or w >= width:
            continue
        cx, cy = rng.uniform(w / 2, width - w / 2), rng.uniform(h / 2, height - h / 2)
        x0, x1 = max(int(cx - w / 2) - 1, 0), min(int(cx + w / 2) + 2, width)
        y0, y1 = max(int(cy - h / 2) - 1, 0), min(int(cy + h / 2) + 2, height)
        m = ((xx[y0:y1, x0:x1] - cx) / (w / 2)) ** 2 + ((yy[y0:y1, x0:x1] - cy) / (h / 2)) ** 2 <= 1.0
        tone = rng.integers(120, 230, 3)
        img[y0:y1, x0:x1][m] = tone
        for ex in (-0.2, 0.2):  # two darker "eyes"
            e = ((xx[y0:y1, x0:x1] - (cx + ex * w)) / (0.08 * w + 0.5)) ** 2 + ((yy[y0:y1, x0:x1] - (cy - 0.1 * h)) / (0.05 * h + 0.5)) ** 2 <= 1.0
            img[y0:y1, x0:x1][e] = tone // 4
        gts.append([cx - w / 2, cy - h / 2, w, h])
    return img, np.array(gts, dtype=np.float64).reshape(-1, 4)


def make_pool_on_device(n: int, height: int, width: int, device, seed: int = 1234, start: int = 0):
    """n synthetic images generated directly in HBM (used by the throughput bench, where thousands of images are
    needed and host-side generation would dominate the run): device-RNG noise background, 3x3 box filter, 12 pasted
    ellipses per image whose parameters come from a seeded host RNG (no device->host syncs)."""
    import torch

    from .ops import ImagePool

    pool = ImagePool(n, height, width, device)
    yy = torch.arange(height, device=device, dtype=torch.float32)[:, None]
    xx = torch.arange(width, device=device, dtype=torch.float32)[None, :]
    g = torch.Generator(device=device)
    chunk = 32
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        g.manual_seed(seed + start + a)
        noise = torch.randint(0, 256, (b - a, 3, height + 2, width + 2), generator=g, device=device, dtype=torch.uint8)
        img = torch.nn.functional.avg_pool2d(noise.float(), 3, stride=1)  # [m,3,H,W]
        for j in range(b - a):
            prm = np.random.default_rng(seed + start + a + j).random((12, 5))
            for f in range(12):
                w = float(np.exp(prm[f, 0] * (np.log(200.0) - np.log(6.0)) + np.log(6.0)))
                h = w * (1.1 + 0.3 * prm[f, 1])
                cx = w / 2 + prm[f, 2] * (width - w)
                cy = h / 2 + prm[f, 3] * max(height - h, 1.0)
                x0, x1 = max(int(cx - w / 2) - 1, 0), min(int(cx + w / 2) + 2, width)
                y0, y1 = max(int(cy - h / 2) - 1, 0), min(int(cy + h / 2) + 2, height)
                m = ((xx[:, x0:x1] - cx) / (w / 2)) ** 2 + ((yy[y0:y1] - cy) / (h / 2)) ** 2 <= 1.0
                region = img[j, :, y0:y1, x0:x1]
                region[:, m] = 120.0 + 110.0 * float(prm[f, 4])
        pool.buf[a:b, :, : width * 3] = img.round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).reshape(b - a, height, width * 3)
    return pool
