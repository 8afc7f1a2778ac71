"""Kernel 1 parity: fused slice gather + letterbox + normalise vs the cv2/torch oracle (bit-exact)."""
import numpy as np
import pytest
import torch

from oracle import letterbox as olb
from oracle import slicing as osl

pytestmark = pytest.mark.gpu


def _oracle_batch(images, entries, src_w, src_h, imgsz, half, reverse):
    outs = []
    for (i, x0, y0) in entries:
        crop = np.ascontiguousarray(images[i][y0:y0 + src_h, x0:x0 + src_w])
        if not reverse:  # ultralytics always flips; feed pre-flipped data to model "no flip"
            crop = np.ascontiguousarray(crop[..., ::-1])
        outs.append(olb.preprocess(crop, imgsz=imgsz, half=half))
    return torch.cat(outs, 0)


CASES = [
    # (H, W, slice_h, slice_w, overlap, imgsz)                      what it exercises
    (768, 1024, 512, 512, 0.2, 1024),     # C2: exact 2x up-scale
    (1080, 1920, 640, 640, 0.2, 1024),    # C1: 1.6x up-scale
    (480, 750, 640, 640, 0.2, 1024),      # image smaller than slice, odd pitch (2250 B rows)
    (1366, 2048, 640, 640, 0.25, 1024),   # real WIDER size
    (333, 517, 200, 300, 0.1, 640),       # odd everything, non-square slices, letterbox padding
    (600, 800, 512, 512, 0.2, 512),       # imgsz == slice: pure copy
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_slices_match_oracle(cuda_device, case, dtype):
    import fsd_b200.ops as ops

    H, W, sh, sw, ov, imgsz = case
    rng = np.random.default_rng(1234)
    images = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(2)]
    boxes = osl.get_slice_bboxes(H, W, sh, sw, overlap_height_ratio=ov, overlap_width_ratio=ov)
    bw, bh = boxes[0][2] - boxes[0][0], boxes[0][3] - boxes[0][1]
    entries = [(i, b[0], b[1]) for i in range(2) for b in boxes]
    pool = ops.ImagePool.from_numpy(images, cuda_device)
    out = ops.gather_letterbox(pool, torch.tensor(entries, dtype=torch.int32), bw, bh, imgsz=imgsz, dtype=dtype)
    ref = _oracle_batch(images, entries, bw, bh, imgsz, dtype == torch.float16, True)
    assert out.shape == ref.shape
    assert torch.equal(out.cpu(), ref), f"max diff {(out.cpu().float() - ref.float()).abs().max()}"


@pytest.mark.parametrize("shape", [(1080, 1920), (768, 1024), (2160, 3840), (1366, 2048), (2340, 4160), (37, 53),
                                   (640, 1280), (768, 1536), (1080, 1920 + 0)])  # 1.25x / 1.5x / 1.875x down: compile-time tap patterns Q = 10 / 12 / 15
def test_full_image_pass_matches_oracle(cuda_device, shape):
    """The perform_standard_pred pass: down-scale (r<1), copy (r==1), exact-2x area path, tiny up-scale."""
    import fsd_b200.ops as ops

    H, W = shape
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    pool = ops.ImagePool.from_numpy([img], cuda_device)
    for dtype in (torch.float16, torch.float32):
        for reverse in (True, False):
            out = ops.gather_letterbox(pool, torch.tensor([[0, 0, 0]], dtype=torch.int32), W, H, imgsz=1024,
                                       dtype=dtype, reverse_channels=reverse)
            ref = _oracle_batch([img], [(0, 0, 0)], W, H, 1024, dtype == torch.float16, reverse)
            assert torch.equal(out.cpu(), ref)
            if dtype == torch.float16:  # channels-last output (what the engine consumes) through the same path
                cl = ops.gather_letterbox(pool, torch.tensor([[0, 0, 0]], dtype=torch.int32), W, H, imgsz=1024, dtype=dtype,
                                          reverse_channels=reverse, channels_last=True)
                assert torch.equal(cl.cpu().contiguous(), ref)


def test_bad_pitch_is_rejected(cuda_device):
    import fsd_b200._cabi as cabi
    import fsd_b200.ops as ops

    pool = ops.ImagePool(1, 64, 64, cuda_device)
    h = cabi.get_handle(0)
    out = torch.empty((1, 3, 64, 64), dtype=torch.float16, device=cuda_device)
    ent = torch.zeros((1, 3), dtype=torch.int32, device=cuda_device)
    rc = h.lib.fsd_gather_letterbox(h.h, pool.buf.data_ptr(), 1, 64, 64, 64 * 3 + 2, 0, ent.data_ptr(), 1, 64, 64,
                                    64, 32, 1, 0, 0, out.data_ptr(), 0)
    assert rc == -2 and b"16-byte" in h.lib.fsd_last_error()


@pytest.mark.parametrize("case", [CASES[0], CASES[1], CASES[4]])
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_channels_last_output(cuda_device, case, dtype):
    """Kernel 1 can emit the network input channels-last (what the cuDNN backbone consumes without a layout copy)."""
    import fsd_b200.ops as ops

    H, W, sh, sw, ov, imgsz = case
    rng = np.random.default_rng(99)
    images = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8)]
    boxes = osl.get_slice_bboxes(H, W, sh, sw, overlap_height_ratio=ov, overlap_width_ratio=ov)
    bw, bh = boxes[0][2] - boxes[0][0], boxes[0][3] - boxes[0][1]
    entries = [(0, b[0], b[1]) for b in boxes]
    pool = ops.ImagePool.from_numpy(images, cuda_device)
    out = ops.gather_letterbox(pool, torch.tensor(entries, dtype=torch.int32), bw, bh, imgsz=imgsz, dtype=dtype, channels_last=True)
    assert out.is_contiguous(memory_format=torch.channels_last)
    ref = _oracle_batch(images, entries, bw, bh, imgsz, dtype == torch.float16, True)
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("size", [(512, 512), (64, 96), (8, 12), (36, 516), (300, 500)])
@pytest.mark.parametrize("channels_last", [False, True])
def test_exact_2x_fast_path_equals_general_kernel_and_oracle(cuda_device, size, channels_last, monkeypatch):
    """imgsz == 2 x slice takes the small-integer fast path (k1_upscale2x.cu); it must equal cv2 AND the general TMA kernel."""
    import fsd_b200.ops as ops

    sh, sw = size
    H, W = sh + 37, sw + 21
    imgsz = 2 * max(sh, sw)
    g = olb.letterbox_geometry(sh, sw, imgsz)
    rng = np.random.default_rng(5)
    images = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(2)]
    entries = [(0, 0, 0), (1, 21, 37), (0, 7, 3), (1, 0, 37)]
    pool = ops.ImagePool.from_numpy(images, cuda_device)
    ent = torch.tensor(entries, dtype=torch.int32)
    fast = ops.gather_letterbox(pool, ent, sw, sh, imgsz=imgsz, dtype=torch.float16, channels_last=channels_last)
    ref = _oracle_batch(images, entries, sw, sh, imgsz, True, True)
    assert torch.equal(fast.cpu(), ref)
    if (g["new_w"], g["new_h"], g["left"], g["top"]) == (2 * sw, 2 * sh, 0, 0):  # the fast path was eligible: cross-check
        monkeypatch.setenv("FSD_K1_GENERIC", "1")
        general = ops.gather_letterbox(pool, ent, sw, sh, imgsz=imgsz, dtype=torch.float16, channels_last=channels_last)
        assert torch.equal(general, fast)


@pytest.mark.parametrize("case", [  # (slice h, slice w, imgsz): ratios 8/5, 4, 1 (copy), 8/5 non-square, 16/5 (NOT sixteenths)
    (640, 640, 1024), (40, 40, 64), (64, 64, 256), (96, 128, 128), (768, 1024, 1024), (160, 200, 320), (20, 40, 128), (100, 100, 320),
    (48, 48, 128), (96, 96, 128), (112, 112, 128), (16, 16, 128), (10, 10, 16), (5, 5, 32)])
@pytest.mark.parametrize("channels_last", [False, True])
def test_sixteenths_path_equals_general_kernel_and_oracle(cuda_device, case, channels_last, monkeypatch):
    """Border-less resizes whose cv2 coefficients are multiples of 128 (ratios 8/5, 4, 1 ...) take the packed 16-bit-lane path
    (k1_sixteenths_kernel); it must equal cv2 (oracle) AND the general TMA kernel bit for bit.  Geometries that do not
    qualify (a letterbox border, 16/5) silently stay on the general kernel — same assertions."""
    import fsd_b200.ops as ops

    sh, sw, imgsz = case
    H, W = sh + 19, sw + 11
    rng = np.random.default_rng(sh * 7 + sw)
    images = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(2)]
    entries = [(0, 0, 0), (1, 11, 19), (0, 5, 2)]
    pool = ops.ImagePool.from_numpy(images, cuda_device)
    ent = torch.tensor(entries, dtype=torch.int32)
    fast = ops.gather_letterbox(pool, ent, sw, sh, imgsz=imgsz, dtype=torch.float16, channels_last=channels_last)
    ref = _oracle_batch(images, entries, sw, sh, imgsz, True, True)
    assert torch.equal(fast.cpu(), ref)
    monkeypatch.setenv("FSD_K1_TABLE16", "1")  # the table-driven variant of the same path (used for non-uniform column patterns)
    table = ops.gather_letterbox(pool, ent, sw, sh, imgsz=imgsz, dtype=torch.float16, channels_last=channels_last)
    assert torch.equal(table, fast)
    monkeypatch.setenv("FSD_K1_GENERIC", "1")
    general = ops.gather_letterbox(pool, ent, sw, sh, imgsz=imgsz, dtype=torch.float16, channels_last=channels_last)
    assert torch.equal(general, fast)


@pytest.mark.parametrize("case", [(768, 1024, 1024, 32), (96, 128, 128, 32), (64, 72, 72, 8), (32, 40, 40, 8), (16, 16, 16, 16)])
@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("reverse", [True, False])
def test_copy_convert_path_equals_sixteenths_path_and_oracle(cuda_device, case, channels_last, reverse, monkeypatch):
    """Scale 1 (the box already has the network-input size): k1_copy_convert_kernel — aligned 16-pixel groups, unaligned x0
    (byte-load path) and a ragged 8-pixel row end — equals cv2 (oracle) and the sixteenths kernel with identity taps bit for bit."""
    import fsd_b200.ops as ops

    sh, sw, imgsz, stride = case
    H, W = sh + 19, sw + 37
    rng = np.random.default_rng(sh * 11 + sw)
    images = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(2)]
    images[0][:2, :3] = [[[0, 255, 1], [254, 128, 127], [2, 3, 253]]] * 2
    entries = [(0, 0, 0), (1, 16, 19), (0, 5, 2), (1, 32, 0), (0, 37, 7)]
    pool = ops.ImagePool.from_numpy(images, cuda_device)
    ent = torch.tensor(entries, dtype=torch.int32)
    kw = dict(imgsz=imgsz, stride=stride, reverse_channels=reverse, dtype=torch.float16, channels_last=channels_last)
    fast = ops.gather_letterbox(pool, ent, sw, sh, **kw)
    assert fast.shape == (5, 3, sh, sw)
    if reverse and stride == 32:
        assert torch.equal(fast.cpu(), _oracle_batch(images, entries, sw, sh, imgsz, True, True))
    want = torch.stack([torch.from_numpy(np.ascontiguousarray(images[i][y:y + sh, x:x + sw, ::-1] if reverse else images[i][y:y + sh, x:x + sw]))
                        for i, x, y in entries]).permute(0, 3, 1, 2).float().div(255.0).half()
    assert torch.equal(fast.cpu(), want)
    monkeypatch.setenv("FSD_K1_NO_COPY", "1")
    assert torch.equal(ops.gather_letterbox(pool, ent, sw, sh, **kw), fast)


def test_in_library_kernel_timing(cuda_device):
    """fsd_kernel_timing_*: one sample per instrumented launch, tagged (kernel id, entries, src_w), positive device time;
    kernels outside the mask are not sampled and re-enabling clears the list."""
    import fsd_b200.ops as ops
    from fsd_b200 import _cabi

    pool = ops.ImagePool(2, 96, 128, cuda_device)
    pool.buf.random_(0, 256)
    ent = torch.tensor([[0, 0, 0], [1, 32, 16], [1, 64, 32]], dtype=torch.int32, device=cuda_device)
    h = _cabi.get_handle(cuda_device.index or 0)
    h.timing_enable((_cabi.FSD_KERNEL_GATHER,))
    ops.gather_letterbox(pool, ent, 64, 64, 128, 32)              # exact-2x fast path
    ops.gather_letterbox(pool, ent[:2], 48, 40, 128, 32)          # general TMA kernel (letterbox border)
    x = torch.randn((1, 16, 8, 8), device=cuda_device).half().contiguous(memory_format=torch.channels_last)
    ops.bias_act(x, torch.zeros(16, device=cuda_device).half(), "silu")  # not in the mask
    got = h.timing_read()
    assert [(k, u, t) for k, u, t, _ in got] == [(_cabi.FSD_KERNEL_GATHER, 3, 64), (_cabi.FSD_KERNEL_GATHER, 2, 48)]
    assert all(0.0 < ms < 50.0 for *_, ms in got)
    h.timing_enable((_cabi.FSD_KERNEL_BIAS_ACT,))
    assert h.timing_read() == []
    ops.bias_act(x, torch.zeros(16, device=cuda_device).half(), "silu")
    (k, units, tag, ms), = h.timing_read()
    assert (k, units, tag) == (_cabi.FSD_KERNEL_BIAS_ACT, 2 * x.numel() * 2, 16) and ms > 0
    h.timing_enable(())
    ops.bias_act(x, torch.zeros(16, device=cuda_device).half(), "silu")
    assert h.timing_read() == []
