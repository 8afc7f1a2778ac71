"""(f4) nvJPEG ingest: a JPEG decoded on the device into the image pool, against PIL's host decode (what the reference's
read_image_as_pil does).  NOT bit-exact by nature — different IDCT / chroma up-sampling implementations — so the bar is a
tolerance: identical size and channel order, mean absolute difference < 1 LSB, 4:4:4 streams within 3 LSB everywhere."""
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


@pytest.mark.parametrize("size", [(768, 1024), (301, 517)])
@pytest.mark.parametrize("subsampling", [0, 2])  # 4:4:4 and 4:2:0
def test_device_decode_tracks_pil(cuda_device, size, subsampling):
    import fsd_b200.ops as ops
    from fsd_b200.synthetic import make_image

    H, W = size
    img, _ = make_image(3, H, W)
    data = _jpeg(img, 92, subsampling)
    assert ops.jpeg_info(data) == (W, H, 3)
    want = np.asarray(Image.open(io.BytesIO(data)).convert("RGB")).astype(np.int32)
    pool = ops.ImagePool(2, H, W, cuda_device)
    pool.upload_jpeg(1, data)
    got = pool.view(1).cpu().numpy().astype(np.int32)
    d = np.abs(got - want)
    assert got.shape == want.shape and d.mean() < 1.0, d.mean()
    if subsampling == 0:
        assert d.max() <= 3, d.max()
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
