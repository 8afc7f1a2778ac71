"""The hand-written kernels AROUND the PyTorch backbones (conv epilogues K5, stem K6, 1x1 convolution K7, SPPF pooling, neck
up-sample + concat) and the backbones built on them, against plain torch: they are not part of the north-star path (the
convolutions stay cuDNN's), so they live apart from the Kernel 4 tests."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


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
            # fp16 network (hand-written epilogues / stem / 1x1 kernels, cuDNN convolutions) vs the fp32 network: relative
            # L2 error per head tensor (fp16 carries ~1e-3 per operation through ~100 layers), and no outlier element
            d = tg.float().cpu() - tw
            rel = (d.norm() / tw.norm().clamp_min(1e-6)).item()
            assert rel < 0.02, (tuple(tw.shape), rel)
            assert d.abs().max().item() < 0.05 * max(1.0, tw.abs().max().item()), d.abs().max().item()


@pytest.mark.parametrize("act", ["silu", "none"])
def test_bias_act_into_concat_slot_with_residual(cuda_device, act):
    """(a5) general epilogue: writes act(x+bias)+residual into a channel slot of a wider channels-last buffer, copies the
    trailing channels to a second destination; equals the torch sequence conv-bias-act, `x + y`, torch.cat exactly."""
    import fsd_b200.ops as ops

    g = torch.Generator().manual_seed(7)
    cl = lambda t: t.half().to(cuda_device).contiguous(memory_format=torch.channels_last)  # noqa: E731
    x = cl(torch.randn((3, 32, 9, 13), generator=g) * 3)
    bias = torch.randn((32,), generator=g).half().to(cuda_device)
    left = cl(torch.randn((3, 16, 9, 13), generator=g))
    resbuf = cl(torch.randn((3, 48, 9, 13), generator=g))
    res = resbuf[:, 8:40]  # a strided slot as residual
    f = torch.nn.functional.silu if act == "silu" else (lambda t: t)
    y = f((x.float() + bias.float().view(1, -1, 1, 1))).half()  # what bias_act_ produces (rounded once)
    assert torch.equal(ops.bias_act_(x.clone(memory_format=torch.channels_last), bias, act), y)
    want = torch.cat((left, res + y), 1)
    buf = torch.full((3, 48, 9, 13), float("nan"), dtype=torch.float16, device=cuda_device).contiguous(memory_format=torch.channels_last)
    buf[:, :16].copy_(left)
    tail = torch.empty((3, 8, 9, 13), dtype=torch.float16, device=cuda_device).contiguous(memory_format=torch.channels_last)
    out = ops.bias_act(x, bias, act, out=buf[:, 16:], residual=res, out2=tail)
    assert out.data_ptr() == buf[:, 16:].data_ptr()
    assert torch.equal(buf, want)
    assert torch.equal(tail, want[:, 40:])
    # in place, no extras == the dedicated in-place kernel
    assert torch.equal(ops.bias_act(x.clone(memory_format=torch.channels_last), bias, act), y)
    # third destination: the result up-sampled 2x (nearest) into a slot of the FPN's next concat buffer
    upbuf = torch.full((3, 40, 18, 26), float("nan"), dtype=torch.float16, device=cuda_device).contiguous(memory_format=torch.channels_last)
    dense = ops.bias_act(x.clone(memory_format=torch.channels_last), bias, act, up2=upbuf[:, :32])
    assert torch.equal(dense, y) and torch.equal(upbuf[:, :32], torch.nn.functional.interpolate(y, scale_factor=2.0, mode="nearest"))
    assert torch.isnan(upbuf[:, 32:]).all()
    with pytest.raises(Exception):
        ops.bias_act(x, bias, act, out=buf[:, 16:].contiguous())  # NCHW-dense is not a channels-last slot
    with pytest.raises(Exception):
        ops.bias_act(x, bias, act, out=buf[:, 4:36])  # slot not 16-byte aligned


@pytest.mark.parametrize("hw", [(32, 32), (24, 32), (3, 5), (1, 1), (40, 40)])
def test_sppf_pool_equals_cascaded_maxpool(cuda_device, hw):
    """(a5) one-launch SPPF pooling == three cascaded MaxPool2d(5,1,2) + torch.cat (bit-exact: max only selects)."""
    import fsd_b200.ops as ops

    g = torch.Generator().manual_seed(hw[0] * 100 + hw[1])
    y = torch.randn((3, 16, *hw), generator=g).half().to(cuda_device)
    m = torch.nn.MaxPool2d(5, 1, 2)
    y1 = m(y); y2 = m(y1); y3 = m(y2)
    want = torch.cat((y, y1, y2, y3), 1)
    buf = torch.zeros((3, 64, *hw), dtype=torch.float16, device=cuda_device).contiguous(memory_format=torch.channels_last)
    buf[:, :16].copy_(y)
    ops.sppf_pool_(buf)
    assert torch.equal(buf, want)


@pytest.mark.parametrize("shape", [(3, 64, 128), (2, 96, 160), (1, 34, 70), (2, 768, 1024)])
def test_stem_conv_matches_torch(cuda_device, shape):
    """(a5) fsd_stem_conv == SiLU(conv2d(x, w, b, stride 2, pad 1)) evaluated in fp32 on the same fp16 inputs; the kernel
    accumulates in fp32 on tensor cores and rounds once, so it must sit within one fp16 rounding of the fp32 result."""
    import fsd_b200.ops as ops

    E, H, W = shape
    g = torch.Generator().manual_seed(H + W)
    x = torch.rand((E, 3, H, W), generator=g).half().to(cuda_device).contiguous(memory_format=torch.channels_last)
    w = (torch.randn((16, 3, 3, 3), generator=g) * 0.4).half().to(cuda_device)
    b = torch.randn((16,), generator=g).half().to(cuda_device)
    got = ops.stem_conv(x, w, b)
    ref = torch.nn.functional.silu(torch.nn.functional.conv2d(x.float(), w.float(), b.float(), stride=2, padding=1))
    assert got.shape == ref.shape and got.is_contiguous(memory_format=torch.channels_last)
    err = (got.float() - ref).abs()
    assert float((err - 1e-3 * ref.abs()).max()) <= 1e-3, float(err.max())
    with pytest.raises(Exception):
        ops.stem_conv(x.contiguous(), w, b)  # NCHW input is rejected, not silently re-laid-out


@pytest.mark.parametrize("kn", [(32, 32), (48, 64), (64, 64), (96, 128), (112, 32), (64, 16), (16, 128), (128, 128)])
@pytest.mark.parametrize("act", ["silu", "none"])
@pytest.mark.parametrize("tensor_cores", [True, False])
def test_pointwise_conv_matches_torch(cuda_device, kn, act, tensor_cores, monkeypatch):
    """(a5) fsd_pointwise_conv == act(conv2d 1x1 + bias) (+ residual) computed in fp32 on the same fp16 inputs, written into
    a concat slot with the trailing channels copied to a second destination; ragged pixel count (not a multiple of 32)."""
    import fsd_b200.ops as ops

    if not tensor_cores:
        monkeypatch.setenv("FSD_K7_NO_TC", "1")  # the mma.sync kernel (Kernel 7); default: tcgen05 (Kernel 10)
    K, N = kn
    g = torch.Generator().manual_seed(K * 1000 + N)
    cl = lambda t: t.half().to(cuda_device).contiguous(memory_format=torch.channels_last)  # noqa: E731
    B, H, W = 3, 37, 29  # 3219 pixels
    xbuf = cl(torch.randn((B, K + 16, H, W), generator=g))
    x = xbuf[:, 8:8 + K]  # the input itself is a channel slot
    w = (torch.randn((N, K, 1, 1), generator=g) / K ** 0.5).half().to(cuda_device)
    bias = torch.randn((N,), generator=g).half().to(cuda_device)
    res = cl(torch.randn((B, N, H, W), generator=g))
    f = torch.nn.functional.silu if act == "silu" else (lambda t: t)
    y = f(torch.nn.functional.conv2d(x.float(), w.float(), bias.float()))
    want = y.half() + res  # act rounded to fp16, then the fp16 residual add — the order torch uses
    buf = torch.full((B, N + 8, H, W), float("nan"), dtype=torch.float16, device=cuda_device).contiguous(memory_format=torch.channels_last)
    tail = torch.empty((B, 8, H, W), dtype=torch.float16, device=cuda_device).contiguous(memory_format=torch.channels_last)
    out = ops.pointwise_conv(x, w, bias, act, out=buf[:, 8:], residual=res, out2=tail)
    assert out.data_ptr() == buf[:, 8:].data_ptr() and torch.isnan(buf[:, :8]).all()
    err = (out.float() - want.float()).abs()
    assert float((err - 2e-3 * want.float().abs()).max()) <= 2e-3, float(err.max())
    assert torch.equal(tail, out[:, N - 8:])
    plain = ops.pointwise_conv(x, w, bias, act)
    assert plain.is_contiguous(memory_format=torch.channels_last)
    err = (plain.float() - y).abs()
    assert float((err - 1e-3 * y.abs()).max()) <= 1e-3, float(err.max())
    assert not ops.pointwise_conv_supported(512, 256) and not ops.pointwise_conv_supported(24, 64)
    assert ops.pointwise_conv_supported(256, 64) == tensor_cores


@pytest.mark.parametrize("kn", [(16, 16), (32, 48), (64, 256), (96, 64), (192, 128), (256, 64), (384, 128), (128, 256), (320, 144), (512, 80)])
@pytest.mark.parametrize("act", ["silu", "none", "lrelu"])
def test_tensor_core_pointwise_conv_wide_shapes(cuda_device, kn, act):
    """(a5) the tcgen05 path of fsd_pointwise_conv on the shapes Kernel 7 cannot take (K up to 512, N up to 256, every slab width
    64 / 32 / 16 and multi-chunk epilogues): fp32 reference on the same fp16 inputs; many tiles per CTA (ring wrap, both accumulators),
    a ragged last tile, slot input/output, residual and second destination."""
    import fsd_b200.ops as ops

    K, N = kn
    g = torch.Generator().manual_seed(K * 7 + N)
    cl = lambda t: t.half().to(cuda_device).contiguous(memory_format=torch.channels_last)  # noqa: E731
    B, H, W = 5, 131, 127  # 83185 pixels = 650 tiles (4-5 per CTA), last tile ragged
    xbuf = cl(torch.randn((B, K + 8, H, W), generator=g))
    x = xbuf[:, 8:]
    w = (torch.randn((N, K, 1, 1), generator=g) / K ** 0.5).half().to(cuda_device)
    bias = torch.randn((N,), generator=g).half().to(cuda_device)
    res = cl(torch.randn((B, N, H, W), generator=g))
    f = {"silu": torch.nn.functional.silu, "none": lambda t: t, "lrelu": lambda t: torch.nn.functional.leaky_relu(t, 0.2)}[act]
    y = f(torch.nn.functional.conv2d(x.float(), w.float(), bias.float()))
    want = y.half() + res
    buf = torch.full((B, N + 24, H, W), float("nan"), dtype=torch.float16, device=cuda_device).contiguous(memory_format=torch.channels_last)
    tail = torch.empty((B, 16, H, W), dtype=torch.float16, device=cuda_device).contiguous(memory_format=torch.channels_last)
    out = ops.pointwise_conv(x, w, bias, act, out=buf[:, 8:8 + N], residual=res, out2=tail)
    assert torch.isnan(buf[:, :8]).all() and torch.isnan(buf[:, 8 + N:]).all()
    err = (out.float() - want.float()).abs()
    assert float((err - 3e-3 * want.float().abs()).max()) <= 3e-3, float(err.max())
    assert torch.equal(tail, out[:, N - 16:])
    plain = ops.pointwise_conv(x, w, bias, act)
    err = (plain.float() - y).abs()
    assert float((err - 2e-3 * y.abs()).max()) <= 2e-3, float(err.max())
    # tiny problem: fewer tiles than SMs, one ragged tile
    xs = cl(torch.randn((1, K, 3, 5), generator=g))
    ys = f(torch.nn.functional.conv2d(xs.float(), w.float(), bias.float()))
    err = (ops.pointwise_conv(xs, w, bias, act).float() - ys).abs()
    assert float((err - 2e-3 * ys.abs()).max()) <= 2e-3, float(err.max())


@pytest.mark.parametrize("kn", [(16, 8), (16, 16), (16, 32), (32, 16), (32, 32), (32, 64), (64, 32), (64, 64), (64, 16), (128, 16), (256, 16),
                                (48, 48), (96, 32)])
@pytest.mark.parametrize("act", ["silu", "none"])
def test_tensor_core_conv3x3_matches_torch(cuda_device, kn, act):
    """(a5) fsd_conv3x3 (implicit GEMM on tcgen05: nine shifted TMA boxes per 16x8 tile, zero fill = padding) == act(conv2d 3x3 pad 1 +
    bias) (+ residual) in fp32 on the same fp16 inputs: ragged image sizes (partial tiles in both directions, an image smaller than a
    tile), several images, slot input / output, every slab width."""
    import fsd_b200.ops as ops

    K, N = kn
    assert ops.conv3x3_supported(K, N) and not ops.conv3x3_supported(24, 16) and not ops.conv3x3_supported(128, 128)
    g = torch.Generator().manual_seed(K * 31 + N)
    cl = lambda t: t.half().to(cuda_device).contiguous(memory_format=torch.channels_last)  # noqa: E731
    f = torch.nn.functional.silu if act == "silu" else (lambda t: t)
    w = (torch.randn((N, K, 3, 3), generator=g) / (3 * K ** 0.5)).half().to(cuda_device)
    bias = torch.randn((N,), generator=g).half().to(cuda_device)
    taps = ops.conv3x3_tap_major(w)
    for B, H, W in ((3, 37, 29), (2, 64, 80), (1, 5, 7), (5, 131, 127)):
        xbuf = cl(torch.randn((B, K + 8, H, W), generator=g))
        x = xbuf[:, 8:]
        res = cl(torch.randn((B, N, H, W), generator=g))
        y = f(torch.nn.functional.conv2d(x.float(), w.float(), bias.float(), padding=1))
        want = y.half() + res
        buf = torch.full((B, N + 16, H, W), float("nan"), dtype=torch.float16, device=cuda_device).contiguous(memory_format=torch.channels_last)
        out = ops.conv3x3(x, taps, bias, act, out=buf[:, 8:8 + N], residual=res)
        assert torch.isnan(buf[:, :8]).all() and torch.isnan(buf[:, 8 + N:]).all()
        err = (out.float() - want.float()).abs()
        assert float((err - 3e-3 * want.float().abs()).max()) <= 3e-3, (B, H, W, float(err.max()))
        plain = ops.conv3x3(x, taps, bias, act)
        assert plain.is_contiguous(memory_format=torch.channels_last)
        err = (plain.float() - y).abs()
        assert float((err - 2e-3 * y.abs()).max()) <= 2e-3, (B, H, W, float(err.max()))


@pytest.mark.parametrize("kn", [(64, 32), (16, 16), (32, 64), (64, 128)])
def test_tensor_core_conv2x2_matches_torch(cuda_device, kn):
    """(a5) fsd_conv2x2 (layer 1 on the space-to-depth stem output) == silu(conv2d 2x2, no padding + bias) in fp32 on the same inputs."""
    import fsd_b200.ops as ops

    K, N = kn
    g = torch.Generator().manual_seed(K + N)
    cl = lambda t: t.half().to(cuda_device).contiguous(memory_format=torch.channels_last)  # noqa: E731
    w = (torch.randn((N, K, 2, 2), generator=g) / (2 * K ** 0.5)).half().to(cuda_device)
    bias = torch.randn((N,), generator=g).half().to(cuda_device)
    taps = w.permute(2, 3, 0, 1).contiguous()
    for B, H, W in ((3, 38, 30), (2, 65, 81), (1, 2, 2), (4, 129, 129)):
        x = cl(torch.randn((B, K, H, W), generator=g))
        y = torch.nn.functional.silu(torch.nn.functional.conv2d(x.float(), w.float(), bias.float()))
        out = ops.conv2x2(x, taps, bias, "silu")
        assert out.shape == (B, N, H - 1, W - 1) and out.is_contiguous(memory_format=torch.channels_last)
        err = (out.float() - y).abs()
        assert float((err - 2e-3 * y.abs()).max()) <= 2e-3, (B, H, W, float(err.max()))


@pytest.mark.parametrize("c", [8, 64, 128, 256])
@pytest.mark.parametrize("act", ["silu", "none"])
def test_depthwise_conv3x3_matches_torch(cuda_device, c, act):
    """(a5) fsd_dwconv3x3 == act(conv2d(groups = C, 3x3, pad 1) + bias) in fp32 on the same fp16 inputs; slot input / output, borders."""
    import fsd_b200.ops as ops

    g = torch.Generator().manual_seed(c)
    cl = lambda t: t.half().to(cuda_device).contiguous(memory_format=torch.channels_last)  # noqa: E731
    f = torch.nn.functional.silu if act == "silu" else (lambda t: t)
    w = (torch.randn((c, 1, 3, 3), generator=g) / 3).half().to(cuda_device)
    bias = torch.randn((c,), generator=g).half().to(cuda_device)
    taps = ops.dwconv3x3_tap_major(w)
    for B, H, W in ((3, 37, 29), (1, 1, 1), (2, 64, 80), (1, 3, 130)):
        xbuf = cl(torch.randn((B, c + 8, H, W), generator=g))
        x = xbuf[:, 8:]
        y = f(torch.nn.functional.conv2d(x.float(), w.float(), bias.float(), padding=1, groups=c))
        buf = torch.full((B, c + 16, H, W), float("nan"), dtype=torch.float16, device=cuda_device).contiguous(memory_format=torch.channels_last)
        out = ops.dwconv3x3(x, taps, bias, act, out=buf[:, 8:8 + c])
        assert torch.isnan(buf[:, :8]).all() and torch.isnan(buf[:, 8 + c:]).all()
        err = (out.float() - y).abs()
        assert float((err - 1e-3 * y.abs()).max()) <= 1e-3, (B, H, W, float(err.max()))
        assert torch.equal(ops.dwconv3x3(x, taps, bias, act), out)


def test_backbone_tensor_core_layers_equal_library_layers(cuda_device, monkeypatch):
    """The YOLO11n-pose graph with the tcgen05 1x1 / 3x3 kernels vs the same weights on cuDNN + fsd_bias_act (+ the mma.sync 1x1 kernel):
    per-level relative error of the raw head tensors stays at fp16 accumulation noise."""
    from fsd_b200.backbones import yolo11_pose as yp

    g = torch.Generator().manual_seed(11)
    model = yp.build_yolo11n_pose().half().to(cuda_device).to(memory_format=torch.channels_last)
    x = torch.rand((2, 3, 256, 320), generator=g).half().to(cuda_device).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        got = model(x)
        monkeypatch.setenv("FSD_K7_NO_TC", "1")
        monkeypatch.setenv("FSD_NO_CONV3_TC", "1")
        monkeypatch.setenv("FSD_NO_DWCONV", "1")
        want = model(x)
    for lv_got, lv_want in zip(got, want):
        for a, b in zip(lv_got, lv_want):
            rel = float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-6))
            assert rel < 2e-2, rel


def test_space_to_depth_stem_equals_plain_layers(cuda_device):
    """(a5) layers 0+1 through the space-to-depth stem (fsd_stem_conv(space_to_depth) + 2x2 convolution) vs the plain
    pair (fsd_stem_conv + cuDNN 3x3 stride-2 convolution): same products and sums, so the results agree to fp16 rounding of
    the accumulation order; the folded tensor itself is an exact re-arrangement of the plain stem output."""
    import fsd_b200.ops as ops
    from fsd_b200.backbones import yolo11_pose as yp

    g = torch.Generator().manual_seed(3)
    model = yp.build_yolo11n_pose().half().to(cuda_device).to(memory_format=torch.channels_last)
    x = torch.rand((3, 3, 96, 160), generator=g).half().to(cuda_device).contiguous(memory_format=torch.channels_last)
    w0 = model.b0.conv.weight.detach().contiguous().clone()
    plain = ops.stem_conv(x, w0, model.b0.conv.bias)
    folded = ops.stem_conv(x, w0, model.b0.conv.bias, space_to_depth=True)
    assert folded.shape == (3, 64, 25, 41) and float(folded[:, :, 0].abs().max()) == 0 and float(folded[:, :, :, 0].abs().max()) == 0
    for dy in range(2):
        for dx in range(2):
            assert torch.equal(folded[:, (dy * 2 + dx) * 16:(dy * 2 + dx + 1) * 16, 1:, 1:], plain[:, :, dy::2, dx::2])
    with torch.no_grad():
        got = model._stem(x)
        old = yp.USE_S2D_STEM
        yp.USE_S2D_STEM = False
        try:
            want = model._stem(x)
        finally:
            yp.USE_S2D_STEM = old
    assert got.shape == want.shape == (3, 32, 24, 40)
    err = (got.float() - want.float()).abs()
    assert float((err - 4e-3 * want.float().abs()).max()) <= 4e-3, float(err.max())
    with pytest.raises(Exception):
        ops.stem_conv(x[:, :, :90], w0, model.b0.conv.bias, space_to_depth=True)  # 45 output rows: not foldable


@pytest.mark.parametrize("scale", [2, 4])
def test_rrdbnet_fast_inference_path(cuda_device, scale):
    """RRDBNet's GPU fp16 path (channels-last convolutions + fsd_bias_act_inplace LeakyReLU epilogue, fused residual scaling) vs
    the plain eager module with the same weights: same network, results agree to fp16 accumulation noise; and vs fp32 on the CPU."""
    from fsd_b200.backbones import rrdbnet

    torch.manual_seed(scale)
    net = rrdbnet.RRDBNet(scale=scale, num_block=3).eval()
    x = torch.rand((2, 3, 40, 56))
    with torch.no_grad():
        want32 = net(x)
        gpu = rrdbnet.RRDBNet(scale=scale, num_block=3).eval()
        gpu.load_state_dict(net.state_dict())
        gpu = gpu.half().to(cuda_device)
        xg = x.half().to(cuda_device)
        rrdbnet.FAST_INFERENCE = False
        try:
            plain = gpu(xg)
        finally:
            rrdbnet.FAST_INFERENCE = True
        fast = gpu(xg)
    assert fast.shape == plain.shape == (2, 3, 40 * scale, 56 * scale) and fast.is_contiguous()
    scale_ref = float(want32.abs().max())
    assert float((fast.float() - plain.float()).abs().max()) <= 2e-2 * scale_ref
    assert float((fast.float().cpu() - want32).abs().max()) <= 3e-2 * scale_ref
