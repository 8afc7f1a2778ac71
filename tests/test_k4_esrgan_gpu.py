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


@pytest.mark.parametrize("act", ["silu", "lrelu", "none"])
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_bias_act_epilogue(cuda_device, act, dtype):
    """(a5) fused conv epilogue vs torch: x + bias then SiLU / LeakyReLU(0.2), channels-last, in place."""
    import fsd_b200.ops as ops

    g = torch.Generator().manual_seed(4)
    x = (torch.randn((3, 32, 17, 23), generator=g) * 3).to(dtype).to(cuda_device).contiguous(memory_format=torch.channels_last)
    b = torch.randn((32,), generator=g).to(dtype).to(cuda_device)
    ref = (x.float() + b.float().view(1, -1, 1, 1))
    ref = {"silu": torch.nn.functional.silu, "lrelu": lambda t: torch.nn.functional.leaky_relu(t, 0.2), "none": lambda t: t}[act](ref)
    got = ops.bias_act_(x.clone(memory_format=torch.channels_last), b, act, 0.2)
    tol = 2e-3 if dtype == torch.float16 else 1e-5
    assert torch.allclose(got.float(), ref, atol=tol, rtol=tol)
    with pytest.raises(Exception):
        ops.bias_act_(x.contiguous(), b, act)


def test_upsample_concat_and_split_conv_equal_torch(cuda_device):
    """(a5) the fused neck op and the split-weight convolution are exact re-arrangements of the PyTorch backbone."""
    import fsd_b200.ops as ops
    from fsd_b200.backbones.yolo11_pose import build_yolo11n_pose

    g = torch.Generator().manual_seed(1)
    a = torch.randn((2, 16, 5, 7), generator=g).half().to(cuda_device).contiguous(memory_format=torch.channels_last)
    b = torch.randn((2, 24, 10, 14), generator=g).half().to(cuda_device).contiguous(memory_format=torch.channels_last)
    ref = torch.cat((torch.nn.functional.interpolate(a, scale_factor=2.0, mode="nearest"), b), 1)
    assert torch.equal(ops.upsample2x_concat(a, b), ref)
    # whole backbone: GPU fp16 (fused epilogue, split convs, fused neck) vs the same weights in plain torch fp32 on the CPU
    model = build_yolo11n_pose()
    x = torch.rand((1, 3, 128, 160), generator=g)
    want = model(x)
    gpu = build_yolo11n_pose().half().to(cuda_device).to(memory_format=torch.channels_last)
    with torch.no_grad():
        got = gpu(x.half().to(cuda_device).contiguous(memory_format=torch.channels_last))
    for lw, lg in zip(want, got):
        for tw, tg in zip(lw, lg):
            assert tw.shape == tg.shape
            err = (tg.float().cpu() - tw).abs().max().item()
            assert err < 0.08 * max(1.0, tw.abs().max().item()), err  # fp16 network vs fp32 network
