"""(f4) nvJPEG ingest: a JPEG decoded on the device into the image pool, against PIL's host decode (what the reference's
read_image_as_pil does).  NOT bit-exact by nature.  Measured on B200 (benchmarks/jpeg_probe.py, gpurun_out -> DESIGN §3): 4:4:4 streams
and smooth content differ by 0.5 LSB on average, at most 4 (IDCT rounding); 4:2:0 streams of per-channel NOISE differ by ~5 LSB on
average in the chroma-heavy channels because nvJPEG replicates chroma samples where libjpeg interpolates them ("fancy up-sampling") —
both decoders are then equally far from the uncompressed original.  The bars below state exactly that."""
import io

import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu


def _jpeg(img, quality, subsampling):
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", quality=quality, subsampling=subsampling)
    return buf.getvalue()


def _smooth(h, w):
    yy, xx = np.mgrid[0:h, 0:w]
    return np.stack([xx * 255 // w, yy * 255 // h, (xx + yy) * 255 // (h + w)], -1).astype(np.uint8)


@pytest.mark.parametrize("size", [(768, 1024), (301, 517)])
@pytest.mark.parametrize("subsampling", [0, 2])  # 4:4:4 and 4:2:0
@pytest.mark.parametrize("content", ["noise", "smooth"])
def test_device_decode_tracks_pil(cuda_device, size, subsampling, content):
    import fsd_b200.ops as ops
    from fsd_b200.synthetic import make_image

    H, W = size
    img = make_image(3, H, W)[0] if content == "noise" else _smooth(H, W)
    data = _jpeg(img, 92, subsampling)
    assert ops.jpeg_info(data) == (W, H, 3)
    want = np.asarray(Image.open(io.BytesIO(data)).convert("RGB")).astype(np.int32)
    pool = ops.ImagePool(2, H, W, cuda_device)
    pool.upload_jpeg(1, data)
    got = pool.view(1).cpu().numpy().astype(np.int32)
    d = np.abs(got - want)
    assert got.shape == want.shape
    if subsampling == 0 or content == "smooth":
        assert d.mean() < 1.0 and d.max() <= 6, (d.mean(), d.max())
    else:  # chroma noise at 4:2:0: the decoders interpolate chroma differently; neither is closer to the source image
        assert d.mean() < 8.0, d.mean()
        assert np.abs(got - img).mean() <= np.abs(want - img).mean() + 1.5
        assert np.abs(got[..., ::-1] - want).mean() > 2 * d.mean()  # (and the channel order is right)
    pool.upload_jpeg(0, data, bgr=True)
    assert np.array_equal(pool.view(0).cpu().numpy()[..., ::-1], pool.view(1).cpu().numpy())
    with pytest.raises(Exception):
        pool.upload_jpeg(0, _jpeg(img[:64, :64].copy(), 90, 0))  # size mismatch is an error, not a partial write
    with pytest.raises(Exception):
        ops.jpeg_info(b"not a jpeg")


def test_batch_api_accepts_encoded_jpegs(cuda_device):
    """get_sliced_prediction_batch on encoded JPEGs == the same call on the arrays nvJPEG decodes them to."""
    import fsd_b200.ops as ops
    from fsd_b200.api import get_sliced_prediction_batch
    from fsd_b200.plugins import YOLOv11PoseDetectionModel
    from fsd_b200.synthetic import make_image
    from fsd_b200.yolo import YOLO

    imgs = [make_image(40 + i, 384, 512)[0] for i in range(3)]
    datas = [_jpeg(im, 95, 0) for im in imgs]
    model = YOLOv11PoseDetectionModel(model=YOLO("random-init"), confidence_threshold=0.4, device="cuda:0", image_size=512)
    pool = ops.ImagePool(3, 384, 512, cuda_device)
    for i, d in enumerate(datas):
        pool.upload_jpeg(i, d)
    decoded = [pool.view(i).cpu().numpy().copy() for i in range(3)]
    a = get_sliced_prediction_batch(datas, model, 256, 256, 0.2, 0.2)
    b = get_sliced_prediction_batch(decoded, model, 256, 256, 0.2, 0.2)
    for ra, rb in zip(a, b):
        assert [p.bbox.to_xyxy() for p in ra.object_prediction_list] == [p.bbox.to_xyxy() for p in rb.object_prediction_list]
        assert (ra.image_width, ra.image_height) == (512, 384) and ra.image.size == (512, 384)  # lazy PIL decode of the bytes
