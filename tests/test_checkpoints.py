"""Weight loading: an ultralytics-style pickled checkpoint is read WITHOUT ultralytics and without running
pickled code, Conv+BatchNorm pairs are folded (eps 1e-3) into the fused layers, and anything that cannot be loaded raises.

ultralytics is absent, so the test pickles a look-alike object graph: real torch modules whose classes live in a temporary
`ultralytics.*` module tree with ultralytics' attribute names (conv / bn / cv1 / cv2 / m / attn / ffn / cv2-cv4 / dfl).  The
tree is removed from sys.modules before loading — exactly the situation on a box without ultralytics."""
import os
import sys
import types

import pytest
import torch
import torch.nn as nn


def _fake_ultralytics():
    """A module tree `ultralytics.nn.modules.*` / `ultralytics.nn.tasks` holding look-alike classes (BN NOT fused)."""
    mods = {}
    for name in ("ultralytics", "ultralytics.nn", "ultralytics.nn.modules", "ultralytics.nn.modules.conv",
                 "ultralytics.nn.modules.block", "ultralytics.nn.modules.head", "ultralytics.nn.tasks"):
        mods[name] = types.ModuleType(name)

    def cls(module, name, base=nn.Module):
        def deco(c):
            c.__module__, c.__qualname__, c.__name__ = module, name, name
            setattr(mods[module], name, c)
            return c
        return deco

    @cls("ultralytics.nn.modules.conv", "Conv")
    class Conv(nn.Module):
        def __init__(self, c1, c2, k=1, s=1, g=1, act=True):
            super().__init__()
            self.conv = nn.Conv2d(c1, c2, k, s, k // 2, groups=g, bias=False)
            self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)
            self.act = nn.SiLU() if act else nn.Identity()

    @cls("ultralytics.nn.modules.conv", "DWConv")
    class DWConv(Conv):
        def __init__(self, c1, c2, k=1, s=1, act=True):
            import math
            super().__init__(c1, c2, k, s, g=math.gcd(c1, c2), act=act)

    @cls("ultralytics.nn.modules.conv", "Concat")
    class Concat(nn.Module):
        def __init__(self):
            super().__init__()
            self.d = 1

    @cls("ultralytics.nn.modules.block", "Bottleneck")
    class Bottleneck(nn.Module):
        def __init__(self, c1, c2, shortcut=True, k=(3, 3), e=0.5):
            super().__init__()
            c_ = int(c2 * e)
            self.cv1, self.cv2 = Conv(c1, c_, k[0], 1), Conv(c_, c2, k[1], 1)
            self.add = shortcut and c1 == c2

    @cls("ultralytics.nn.modules.block", "C3k")
    class C3k(nn.Module):
        def __init__(self, c1, c2, n=2, shortcut=True, e=0.5, k=3):
            super().__init__()
            c_ = int(c2 * e)
            self.cv1, self.cv2, self.cv3 = Conv(c1, c_, 1, 1), Conv(c1, c_, 1, 1), Conv(2 * c_, c2, 1)
            self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, k=(k, k), e=1.0) for _ in range(n)))

    @cls("ultralytics.nn.modules.block", "C3k2")
    class C3k2(nn.Module):
        def __init__(self, c1, c2, n=1, c3k=False, e=0.5, shortcut=True):
            super().__init__()
            self.c = int(c2 * e)
            self.cv1, self.cv2 = Conv(c1, 2 * self.c, 1, 1), Conv((2 + n) * self.c, c2, 1)
            self.m = nn.ModuleList(C3k(self.c, self.c, 2, shortcut) if c3k else Bottleneck(self.c, self.c, shortcut) for _ in range(n))

    @cls("ultralytics.nn.modules.block", "SPPF")
    class SPPF(nn.Module):
        def __init__(self, c1, c2, k=5):
            super().__init__()
            self.cv1, self.cv2 = Conv(c1, c1 // 2, 1, 1), Conv(c1 // 2 * 4, c2, 1, 1)
            self.m = nn.MaxPool2d(k, 1, k // 2)

    @cls("ultralytics.nn.modules.block", "Attention")
    class Attention(nn.Module):
        def __init__(self, dim, num_heads=8, attn_ratio=0.5):
            super().__init__()
            self.num_heads, self.head_dim = num_heads, dim // num_heads
            self.key_dim = int(self.head_dim * attn_ratio)
            self.qkv = Conv(dim, dim + self.key_dim * num_heads * 2, 1, act=False)
            self.proj, self.pe = Conv(dim, dim, 1, act=False), Conv(dim, dim, 3, 1, g=dim, act=False)

    @cls("ultralytics.nn.modules.block", "PSABlock")
    class PSABlock(nn.Module):
        def __init__(self, c, attn_ratio=0.5, num_heads=4):
            super().__init__()
            self.attn = Attention(c, num_heads, attn_ratio)
            self.ffn = nn.Sequential(Conv(c, c * 2, 1), Conv(c * 2, c, 1, act=False))

    @cls("ultralytics.nn.modules.block", "C2PSA")
    class C2PSA(nn.Module):
        def __init__(self, c1, c2, n=1, e=0.5):
            super().__init__()
            self.c = int(c1 * e)
            self.cv1, self.cv2 = Conv(c1, 2 * self.c, 1, 1), Conv(2 * self.c, c1, 1)
            self.m = nn.Sequential(*(PSABlock(self.c, 0.5, max(1, self.c // 64)) for _ in range(n)))

    @cls("ultralytics.nn.modules.block", "DFL")
    class DFL(nn.Module):
        def __init__(self, c1=16):
            super().__init__()
            self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)

    @cls("ultralytics.nn.modules.head", "Pose")
    class Pose(nn.Module):
        def __init__(self, nc, kpt_shape, ch):
            super().__init__()
            self.nc, self.kpt_shape, self.reg_max = nc, kpt_shape, 16
            nk = kpt_shape[0] * kpt_shape[1]
            c2, c3, c4 = max(16, ch[0] // 4, 64), max(ch[0], min(nc, 100)), max(ch[0] // 4, nk)
            self.cv2 = nn.ModuleList(nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 64, 1)) for x in ch)
            self.cv3 = nn.ModuleList(nn.Sequential(nn.Sequential(DWConv(x, x, 3), Conv(x, c3, 1)),
                                                   nn.Sequential(DWConv(c3, c3, 3), Conv(c3, c3, 1)), nn.Conv2d(c3, nc, 1)) for x in ch)
            self.cv4 = nn.ModuleList(nn.Sequential(Conv(x, c4, 3), Conv(c4, c4, 3), nn.Conv2d(c4, nk, 1)) for x in ch)
            self.dfl = DFL(16)

    @cls("ultralytics.nn.tasks", "PoseModel")
    class PoseModel(nn.Module):
        def __init__(self, width=0.25, nc=1, kpt_shape=(5, 3)):
            super().__init__()
            ch = lambda c: int(-(-min(c, 1024) * width // 8) * 8)  # noqa: E731
            c64, c128, c256, c512, c1024 = ch(64), ch(128), ch(256), ch(512), ch(1024)
            up = lambda: nn.Upsample(None, 2, "nearest")  # noqa: E731
            self.model = nn.Sequential(
                Conv(3, c64, 3, 2), Conv(c64, c128, 3, 2), C3k2(c128, c256, 1, False, 0.25), Conv(c256, c256, 3, 2),
                C3k2(c256, c512, 1, False, 0.25), Conv(c512, c512, 3, 2), C3k2(c512, c512, 1, True), Conv(c512, c1024, 3, 2),
                C3k2(c1024, c1024, 1, True), SPPF(c1024, c1024, 5), C2PSA(c1024, c1024, 1),
                up(), Concat(), C3k2(c1024 + c512, c512, 1, False), up(), Concat(), C3k2(c512 + c512, c256, 1, False),
                Conv(c256, c256, 3, 2), Concat(), C3k2(c256 + c512, c512, 1, False), Conv(c512, c512, 3, 2), Concat(),
                C3k2(c512 + c1024, c1024, 1, True), Pose(nc, kpt_shape, (c256, c512, c1024)))
            self.names = {0: "face"}

    return mods, PoseModel


@pytest.mark.parametrize("width,scale", [(0.25, "n"), (0.50, "s")])
def test_ultralytics_pickle_is_read_without_ultralytics(tmp_path, width, scale):
    from fsd_b200 import checkpoints as ck
    from fsd_b200.yolo import YOLO

    mods, PoseModel = _fake_ultralytics()
    sys.modules.update(mods)
    try:
        torch.manual_seed(3)
        net = PoseModel(width=width)
        for m in net.modules():  # non-trivial BatchNorm statistics so that folding is really exercised
            if isinstance(m, nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.5); m.running_var.uniform_(0.5, 2.0); m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.3)
        path = str(tmp_path / "best.pt")
        torch.save({"epoch": 3, "model": net.half(), "ema": None, "train_args": {"imgsz": 640}}, path)
        net = net.float().eval()
    finally:
        for k in mods:
            sys.modules.pop(k, None)
    assert "ultralytics" not in sys.modules
    with pytest.raises(Exception):
        torch.load(path, weights_only=True)  # the plain safe loader cannot read it
    yolo = YOLO(path)
    assert yolo.info["scale"] == scale and yolo.info["nc"] == 1 and tuple(yolo.info["kpt_shape"]) == (5, 3)
    ours = yolo.model
    # folded layer == conv -> bn of the pickled graph, layer by layer on random inputs
    pairs = [(net.model[0], ours.b0), (net.model[2].m[0].cv1, ours.b2.m[0].cv1), (net.model[6].m[0].m[1].cv2, ours.b6.m[0].m[1].cv2),
             (net.model[10].m[0].attn.pe, ours.b10.m[0].attn.pe), (net.model[23].cv3[1][0][0], ours.head.cv3[1][0][0]),
             (net.model[23].cv4[2][1], ours.head.cv4[2][1])]
    for ref, mine in pairs:
        x = torch.randn(2, ref.conv.in_channels, 12, 12)
        want = ref.bn(ref.conv(x))
        got = mine.conv(x)
        assert torch.allclose(got, want, atol=2e-3, rtol=2e-3)  # the checkpoint stores fp16 weights
    assert torch.equal(ours.head.cv2[0][2].weight, net.model[23].cv2[0][2].weight)
    # own format round trip keeps the architecture
    p2 = str(tmp_path / "own.pt")
    yolo.save(p2)
    again = YOLO(p2)
    assert again.model.arch == ours.arch
    for (k1, v1), (k2, v2) in zip(ours.state_dict().items(), again.model.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


def test_unloadable_weights_raise(tmp_path):
    from fsd_b200 import checkpoints as ck
    from fsd_b200.yolo import YOLO

    with pytest.raises(FileNotFoundError):
        YOLO(str(tmp_path / "nope.pt"))
    junk = tmp_path / "junk.pt"
    junk.write_bytes(b"not a checkpoint")
    with pytest.raises(Exception):
        YOLO(str(junk))
    # a pickle that references a global outside ultralytics.* / torch.nn.modules.* is refused, not executed
    evil = str(tmp_path / "evil.pt")
    torch.save({"model": os.path.join}, evil)
    with pytest.raises(ck.CheckpointError):
        ck.read_ultralytics_checkpoint(evil)
    assert YOLO(str(tmp_path / "nope.pt"), allow_random_init=True).info["source"] == "random-init"
    with pytest.raises(FileNotFoundError):
        ck.load_rrdbnet_state(str(tmp_path / "RealESRGAN_x4plus.pth"))
    sd = {"params_ema": {"a": torch.ones(1)}, "params": {"a": torch.zeros(1)}}
    torch.save(sd, str(tmp_path / "w.pth"))
    assert float(ck.load_rrdbnet_state(str(tmp_path / "w.pth"))["a"]) == 1.0
