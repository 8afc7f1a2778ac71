"""Out-of-bounds WRITE detection with guard bands (compute-sanitizer is closed on this GPU pool — profiles/r2_sanitizer_unavailable.txt —
so the kernels' stores are fenced by hand): every output buffer is a window inside a larger allocation whose surroundings hold
a sentinel pattern; after the kernel the sentinel must be intact, for the odd shapes, partial tiles and unaligned windows where
an indexing slip would land.  Results inside the window are compared with the oracle by the parity tests; this file is about
the bytes around it."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 4096  # elements on both sides


def guarded(shape, dtype, device, memory_format=torch.contiguous_format, sentinel=None):
    n = int(np.prod(shape))
    raw = torch.empty(n + 2 * GUARD, dtype=dtype, device=device)
    pat = sentinel if sentinel is not None else (-7777.0 if dtype.is_floating_point else 93)
    raw.fill_(pat)
    if memory_format == torch.channels_last:
        b, c, h, w = shape
        win = raw[GUARD:GUARD + n].view(b, h, w, c).permute(0, 3, 1, 2)
    else:
        win = raw[GUARD:GUARD + n].view(shape)

    def intact():
        return bool((raw[:GUARD] == pat).all()) and bool((raw[GUARD + n:] == pat).all())

    return raw, win, intact


@pytest.mark.parametrize("case", [(768, 1024, 512, 512, 1024), (1080, 1920, 640, 640, 1024), (333, 517, 200, 300, 640), (600, 800, 512, 512, 512),
                                  (480, 750, 640, 640, 1024), (1080, 1920, 1080, 1920, 1024), (640, 1280, 640, 1280, 1024)])
@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_k1_writes_stay_inside_the_network_input(cuda_device, case, channels_last, dtype):
    import fsd_b200._cabi as cabi
    import fsd_b200.ops as ops

    H, W, sh, sw, imgsz = case
    rng = np.random.default_rng(1)
    pool = ops.ImagePool.from_numpy([rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(2)], cuda_device)
    sw, sh = min(sw, W), min(sh, H)
    entries = torch.tensor([[0, 0, 0], [1, W - sw, H - sh], [0, (W - sw) // 2, (H - sh) // 3]], dtype=torch.int32)
    g = cabi.letterbox_geometry(sh, sw, imgsz, 32)
    raw, win, intact = guarded((3, 3, g["out_h"], g["out_w"]), dtype, cuda_device,
                               torch.channels_last if channels_last else torch.contiguous_format)
    out = ops.gather_letterbox(pool, entries, sw, sh, imgsz=imgsz, dtype=dtype, out=win, channels_last=channels_last)
    torch.cuda.synchronize()
    assert intact(), "Kernel 1 wrote outside its output tensor"
    assert bool(torch.isfinite(out.float()).all()) and float(out.float().min()) >= 0.0 and float(out.float().max()) <= 1.0


@pytest.mark.parametrize("conf,cap", [(0.5, 64), (0.01, 16), (0.01, 5000)])
def test_k2a_writes_stay_inside_the_candidate_rows(cuda_device, conf, cap):
    import fsd_b200.ops as ops

    g = torch.Generator().manual_seed(2)
    B, levels = 3, []
    for s in (8, 16, 32):
        h, w = 256 // s, 320 // s
        levels.append(tuple(t.half().to(cuda_device) for t in (torch.randn((B, 64, h, w), generator=g) * 1.5 + 1, torch.randn((B, 1, h, w), generator=g) * 2 - 3,
                                                               torch.randn((B, 15, h, w), generator=g))))
    raw, cand, intact = guarded((B, cap, ops.ROW), torch.float32, cuda_device)
    rawc, count, intact_c = guarded((B,), torch.int32, cuda_device)
    ops.pose_decode(levels, conf, cand=cand, count=count)
    torch.cuda.synchronize()
    assert intact() and intact_c(), "Kernel 2a wrote outside the candidate rows / counters (capacity overflow must drop, not spill)"
    assert int(count.max()) > 0


@pytest.mark.parametrize("size,tile,scale,dtype", [((1080, 1920), 400, 2, torch.float16), ((270, 482), 128, 2, torch.float16), ((61, 77), 32, 4, torch.float16),
                                                   ((61, 77), 32, 2, torch.float32), ((600, 500), 256, 4, torch.float16)])
def test_k4_writes_stay_inside_tiles_and_output(cuda_device, size, tile, scale, dtype):
    import fsd_b200.ops as ops

    H, W = size
    rng = np.random.default_rng(3)
    frames = torch.from_numpy(rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)).to(cuda_device)
    table, _ = ops.esrgan_tile_table(H, W, scale, tile, 10, 0)
    n_in, n_out = ops.esrgan_tile_elems(table), ops.esrgan_tile_elems(table, scale, out=True)
    raw, tiles, intact = guarded((2, n_in), dtype, cuda_device)
    _, tab_dev = ops.esrgan_crop(frames, table, scale, 0, dtype, tiles=tiles)   # fp16 + even sizes: the TMA / bulk-store pipeline
    torch.cuda.synchronize()
    assert intact(), "Kernel 4 crop wrote outside the packed tile buffer"
    outs = torch.rand((2, n_out), device=cuda_device).to(dtype)
    rawo, dst, intact_o = guarded((2, H * scale, W * scale, 3), torch.uint8, cuda_device)
    ops.esrgan_stitch(outs, table, tab_dev, scale, H, W, out=dst)
    torch.cuda.synchronize()
    assert intact_o(), "Kernel 4 stitch wrote outside the output image"
