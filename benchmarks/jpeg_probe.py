"""nvJPEG (fsd_jpeg_decode) vs PIL/libjpeg on the same streams: mean / max absolute difference by content and chroma sub-sampling,
and each decoder's distance to the uncompressed original.  One JSON line per case (sets the bars of tests/test_jpeg_ingest_gpu.py)."""
import io
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from PIL import Image

import fsd_b200  # noqa: F401
from fsd_b200 import ops
from fsd_b200.synthetic import make_image


def smooth(h, w):
    yy, xx = np.mgrid[0:h, 0:w]
    return np.stack([(xx * 255 // w), (yy * 255 // h), ((xx + yy) * 255 // (h + w))], -1).astype(np.uint8)


dev = torch.device("cuda:0")
for name, img in (("noise", make_image(3, 768, 1024)[0]), ("smooth", smooth(768, 1024)), ("gray-noise", np.repeat(make_image(3, 768, 1024)[0][..., :1], 3, 2))):
    for q in (75, 92, 100):
        for sub in (0, 2):
            buf = io.BytesIO()
            Image.fromarray(img).save(buf, format="JPEG", quality=q, subsampling=sub)
            data = buf.getvalue()
            want = np.asarray(Image.open(io.BytesIO(data)).convert("RGB")).astype(np.int32)
            pool = ops.ImagePool(1, 768, 1024, dev)
            pool.upload_jpeg(0, data)
            got = pool.view(0).cpu().numpy().astype(np.int32)
            d = np.abs(got - want)
            print(json.dumps({"content": name, "quality": q, "subsampling": sub, "mean": float(d.mean()), "max": int(d.max()),
                              "per_channel_mean": [float(d[..., c].mean()) for c in range(3)],
                              "nvjpeg_vs_orig": float(np.abs(got - img).mean()), "pil_vs_orig": float(np.abs(want - img).mean()),
                              "swapped_mean": float(np.abs(got[..., ::-1] - want).mean())}), flush=True)
