"""Kernel 4 parity: Real-ESRGAN tile crop / stitch vs the RealESRGANer restatement (oracle/esrgan.py). Bit-exact."""
import numpy as np
import pytest
import torch

from oracle.esrgan import NearestUpsampler, RealESRGANer

pytestmark = pytest.mark.gpu


class _Affine(torch.nn.Module):
    """nearest up-sample then an affine map that leaves [0,1] on both sides (exercises clamp + rounding)."""

    def __init__(self, scale):
        super().__init__()
        self.scale = scale

    def forward(self, x):
        return torch.nn.functional.interpolate(x.float(), scale_factor=self.scale, mode="nearest").to(x.dtype) * 1.25 - 0.125


SIZES = [(1080, 1920), (1079, 1919), (37, 53), (600, 500), (256, 256), (5, 7)]


@pytest.mark.parametrize("size", SIZES)
@pytest.mark.parametrize("tile", [256, 400, 200])
@pytest.mark.parametrize("scale", [2, 4])
def test_crop_and_stitch_match_oracle(cuda_device, size, tile, scale):
    import fsd_b200.ops as ops

    H, W = size
    rng = np.random.default_rng(H * 7 + W)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    for half, model in ((True, _Affine(scale)), (False, NearestUpsampler(scale))):
        ref = RealESRGANer(scale=scale, model=model, tile=tile, tile_pad=10, pre_pad=0, half=half)
        want, _ = ref.enhance(img, outscale=scale)
        table, (ph, pw) = ops.esrgan_tile_table(H, W, scale, tile, 10, 0)
        assert (ph, pw) == tuple(ref.img.shape[2:])
        assert len(table) == len(ref.last_tiles)
        dimg = torch.from_numpy(img).to(cuda_device)
        dtype = torch.float16 if half else torch.float32
        tiles, tab_dev = ops.esrgan_crop(dimg, table, scale, 0, dtype)
        outbuf = ops.esrgan_out_buffer(table, scale, dtype, cuda_device)
        for i, row in enumerate(table):
            t_in = ops.tile_view(tiles, row)
            assert torch.equal(t_in.cpu(), ref.last_tiles[i][0]), f"tile {i} crop differs"
            ops.tile_view(outbuf, row, scale, out=True).copy_(model(t_in))
        got = ops.esrgan_stitch(outbuf, table, tab_dev, scale, H, W)
        assert got.shape == want.shape
        assert np.array_equal(got.cpu().numpy(), want)
        if not half:  # identity up-sampler: the stitched image is the nearest-neighbour up-scaled input, exactly
            assert np.array_equal(want, np.repeat(np.repeat(img, scale, 0), scale, 1))


def test_pre_pad_reflect(cuda_device):
    import fsd_b200.ops as ops

    H, W, scale, tile, pre = 61, 45, 2, 32, 10
    img = np.random.default_rng(1).integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref = RealESRGANer(scale=scale, model=NearestUpsampler(scale), tile=tile, tile_pad=10, pre_pad=pre, half=False)
    want, _ = ref.enhance(img, outscale=scale)
    table, _ = ops.esrgan_tile_table(H, W, scale, tile, 10, pre)
    dimg = torch.from_numpy(img).to(cuda_device)
    tiles, tab_dev = ops.esrgan_crop(dimg, table, scale, pre, torch.float32)
    outbuf = ops.esrgan_out_buffer(table, scale, torch.float32, cuda_device)
    for i, row in enumerate(table):
        t_in = ops.tile_view(tiles, row)
        assert torch.equal(t_in.cpu(), ref.last_tiles[i][0])
        ops.tile_view(outbuf, row, scale, out=True).copy_(NearestUpsampler(scale)(t_in))
    got = ops.esrgan_stitch(outbuf, table, tab_dev, scale, H, W)
    assert np.array_equal(got.cpu().numpy(), want)


def test_round_half_even_and_clamp(cuda_device):
    import fsd_b200.ops as ops

    H, W, scale = 16, 32, 1
    table, _ = ops.esrgan_tile_table(H, W, 4, 64, 10, 0)  # one tile, scale handled manually below
    vals = np.concatenate([(np.arange(256, dtype=np.float32) + 0.5) / 255.0, np.linspace(-0.5, 1.5, 256, dtype=np.float32)])
    plane = np.resize(vals, (H * 4, W * 4)).astype(np.float32)
    out = torch.from_numpy(np.stack([plane, plane[::-1].copy(), plane.T.reshape(H * 4, W * 4)])).to(cuda_device)
    buf = ops.esrgan_out_buffer(table, 4, torch.float32, cuda_device)
    ops.tile_view(buf, table[0], 4, out=True).copy_(out[None])
    got = ops.esrgan_stitch(buf, table, torch.from_numpy(table).to(cuda_device), 4, H, W).cpu().numpy()
    ref = out.cpu().clamp_(0, 1).numpy()
    ref = (np.transpose(ref[[2, 1, 0]], (1, 2, 0)) * 255.0).round().astype(np.uint8)
    assert np.array_equal(got, ref)




def test_batched_crop_equals_per_frame_and_tma_equals_load_store(cuda_device, monkeypatch):
    """fp16 tiles of even-sized frames take the bulk-copy pipeline (TMA boxes in, bulk stores out: k4_crop_tma.cu); it must
    write exactly what the load/store kernel writes, frame by frame and for a batch of frames in one launch."""
    import fsd_b200.ops as ops

    rng = np.random.default_rng(3)
    for (H, W, tile) in ((1080, 1920, 400), (600, 500, 256), (270, 482, 128), (64, 96, 400)):
        frames = torch.from_numpy(rng.integers(0, 256, (3, H, W, 3), dtype=np.uint8)).to(cuda_device)
        table, _ = ops.esrgan_tile_table(H, W, 2, tile, 10, 0)
        batch, _ = ops.esrgan_crop(frames, table, 2, 0, torch.float16)
        singles = [ops.esrgan_crop(frames[i], table, 2, 0, torch.float16)[0] for i in range(3)]
        monkeypatch.setenv("FSD_K4_NO_TMA", "1")
        plain, _ = ops.esrgan_crop(frames, table, 2, 0, torch.float16)
        monkeypatch.delenv("FSD_K4_NO_TMA")
        for i in range(3):
            for row in table:  # (the 0-7 padding elements between packed tiles are never written: compare tile views)
                a = ops.tile_view(batch[i], row)
                assert torch.equal(a, ops.tile_view(singles[i], row)) and torch.equal(a, ops.tile_view(plain[i], row))
                assert float(a.float().max()) <= 1.0 and float(a.float().min()) >= 0.0


@pytest.mark.parametrize("size,tile,scale", [((61, 77), 32, 2), ((130, 203), 64, 4), ((97, 64), 200, 2)])
def test_batched_crop_stitch_equals_per_image(cuda_device, size, tile, scale):
    """K4 batch mode (n_images > 1, one launch) is bit-identical to the reference's frame-by-frame calls, including a
    batch taken as a strided view (image pitch > H*W*3) and the product-level RealESRGANer.enhance_device."""
    import fsd_b200.ops as ops
    from fsd_b200.enhancer import RealESRGANer as GpuESRGANer

    H, W = size
    rng = np.random.default_rng(H + W + scale)
    imgs = torch.from_numpy(rng.integers(0, 256, (5, H + 3, W, 3), dtype=np.uint8)).to(cuda_device)
    batch = imgs[:, :H]  # image pitch (H+3)*W*3 != H*W*3
    table, _ = ops.esrgan_tile_table(H, W, scale, tile, 10, 0)
    tiles, tab_dev = ops.esrgan_crop(batch, table, scale, 0, torch.float16)
    assert tiles.shape[0] == 5
    outs = ops.esrgan_out_buffer(table, scale, torch.float16, cuda_device, n_images=5)
    outs.copy_(torch.rand(outs.shape, device=cuda_device) * 1.2 - 0.1)
    got = ops.esrgan_stitch(outs, table, tab_dev, scale, H, W)
    assert got.shape == (5, H * scale, W * scale, 3)
    for n in range(5):
        t1, _ = ops.esrgan_crop(batch[n].contiguous(), table, scale, 0, torch.float16, tab_dev=tab_dev)
        for row in table:  # (the 0-7 padding elements between packed tiles are never written: compare tile views)
            assert torch.equal(ops.tile_view(t1, row), ops.tile_view(tiles[n], row)), f"image {n}: batched crop differs"
        assert torch.equal(ops.esrgan_stitch(outs[n], table, tab_dev, scale, H, W), got[n]), f"image {n}: batched stitch differs"
    up = GpuESRGANer(scale=scale, model=NearestUpsampler(scale), tile=tile, tile_pad=10, pre_pad=0, half=False, max_tile_batch=7)
    whole = up.enhance_device(batch.contiguous())
    assert torch.equal(whole, torch.repeat_interleave(torch.repeat_interleave(batch, scale, 1), scale, 2))
    assert torch.equal(up.enhance_device(batch[2].contiguous()), whole[2])
