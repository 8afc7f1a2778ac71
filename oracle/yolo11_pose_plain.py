"""Oracle YOLO11-pose network (test infrastructure / CPU baseline, see oracle/__init__): the plain-PyTorch statement of
ultralytics' yolo11-pose.yaml graph ([EXT ultralytics], loaded by the reference at utils/yolo_wrapper.py:55 and run at
:74-80), written the way ultralytics runs it — nn.Conv2d + SiLU, torch.cat, nn.Upsample, nn.MaxPool2d, matmul/softmax
attention — with NO hand-written kernel, concat-slot or layout trick.  `bench.py --impl reference` and the fp16-vs-fp32
AP test build their CPU network from here, so the reference arm imports nothing from the product's backbones.

Parameter names equal the product's (`b0..b10, h13..h22, head.cv2/cv3/cv4`), so a state_dict moves between the two; the
forward returns the raw per-level head tensors [(box [B,64,h,w], cls [B,nc,h,w], kpt [B,nk,h,w])] that
oracle.yolo_head.decode_head consumes.  Conv+BN are kept fused (ultralytics fuses before inference).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


def _div8(x):
    return int(math.ceil(x / 8) * 8)


class Conv(nn.Module):
    def __init__(self, c1, c2, k=1, s=1, g=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2, groups=g, bias=True)
        self.act = nn.SiLU() if act else nn.Identity()

    def forward(self, x):
        return self.act(self.conv(x))


class DWConv(Conv):
    def __init__(self, c1, c2, k=1, s=1, act=True):
        super().__init__(c1, c2, k, s, g=math.gcd(c1, c2), act=act)


class Bottleneck(nn.Module):
    def __init__(self, c1, c2, shortcut=True, k=(3, 3), e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1, self.cv2 = Conv(c1, c_, k[0], 1), Conv(c_, c2, k[1], 1)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        return x + self.cv2(self.cv1(x)) if self.add else self.cv2(self.cv1(x))


class C3k(nn.Module):
    def __init__(self, c1, c2, n=2, shortcut=True, e=0.5, k=3):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1, self.cv2, self.cv3 = Conv(c1, c_, 1, 1), Conv(c1, c_, 1, 1), Conv(2 * c_, c2, 1)
        self.m = nn.ModuleList(Bottleneck(c_, c_, shortcut, k=(k, k), e=1.0) for _ in range(n))

    def forward(self, x):
        y = self.cv1(x)
        for m in self.m:
            y = m(y)
        return self.cv3(torch.cat((y, self.cv2(x)), 1))


class C3k2(nn.Module):
    def __init__(self, c1, c2, n=1, c3k=False, e=0.5, shortcut=True):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1, self.cv2 = Conv(c1, 2 * self.c, 1, 1), Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(C3k(self.c, self.c, 2, shortcut) if c3k else Bottleneck(self.c, self.c, shortcut) for _ in range(n))

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        y.extend(m(y[-1]) for m in self.m)
        return self.cv2(torch.cat(y, 1))


class SPPF(nn.Module):
    def __init__(self, c1, c2, k=5):
        super().__init__()
        self.cv1, self.cv2 = Conv(c1, c1 // 2, 1, 1), Conv(c1 // 2 * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)

    def forward(self, x):
        y = [self.cv1(x)]
        y.extend(self.m(y[-1]) for _ in range(3))
        return self.cv2(torch.cat(y, 1))


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, attn_ratio=0.5):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.key_dim = int(self.head_dim * attn_ratio)
        self.scale = self.key_dim ** -0.5
        self.qkv = Conv(dim, dim + self.key_dim * num_heads * 2, 1, act=False)
        self.proj = Conv(dim, dim, 1, act=False)
        self.pe = Conv(dim, dim, 3, 1, g=dim, act=False)

    def forward(self, x):
        B, C, H, W = x.shape
        N = H * W
        qkv = self.qkv(x)
        q, k, v = qkv.view(B, self.num_heads, self.key_dim * 2 + self.head_dim, N).split([self.key_dim, self.key_dim, self.head_dim], dim=2)
        attn = ((q.transpose(-2, -1) @ k) * self.scale).softmax(dim=-1)
        x = (v @ attn.transpose(-2, -1)).view(B, C, H, W) + self.pe(v.reshape(B, C, H, W))
        return self.proj(x)


class PSABlock(nn.Module):
    def __init__(self, c, attn_ratio=0.5, num_heads=4):
        super().__init__()
        self.attn = Attention(c, num_heads=num_heads, attn_ratio=attn_ratio)
        self.ffn = nn.Sequential(Conv(c, c * 2, 1), Conv(c * 2, c, 1, act=False))

    def forward(self, x):
        x = x + self.attn(x)
        return x + self.ffn(x)


class C2PSA(nn.Module):
    def __init__(self, c1, c2, n=1, e=0.5):
        super().__init__()
        self.c = int(c1 * e)
        self.cv1, self.cv2 = Conv(c1, 2 * self.c, 1, 1), Conv(2 * self.c, c1, 1)
        self.m = nn.ModuleList(PSABlock(self.c, attn_ratio=0.5, num_heads=max(1, self.c // 64)) for _ in range(n))

    def forward(self, x):
        a, b = self.cv1(x).split((self.c, self.c), dim=1)
        for m in self.m:
            b = m(b)
        return self.cv2(torch.cat((a, b), 1))


class PoseHead(nn.Module):
    def __init__(self, nc, kpt_shape, ch):
        super().__init__()
        self.nc, self.nk = nc, kpt_shape[0] * kpt_shape[1]
        c2, c3, c4 = max(16, ch[0] // 4, 64), max(ch[0], min(nc, 100)), max(ch[0] // 4, self.nk)
        self.cv2 = nn.ModuleList(nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 64, 1)) for x in ch)
        self.cv3 = nn.ModuleList(nn.Sequential(nn.Sequential(DWConv(x, x, 3), Conv(x, c3, 1)),
                                               nn.Sequential(DWConv(c3, c3, 3), Conv(c3, c3, 1)), nn.Conv2d(c3, nc, 1)) for x in ch)
        self.cv4 = nn.ModuleList(nn.Sequential(Conv(x, c4, 3), Conv(c4, c4, 3), nn.Conv2d(c4, self.nk, 1)) for x in ch)

    def forward(self, feats):
        return [(self.cv2[i](x), self.cv3[i](x), self.cv4[i](x)) for i, x in enumerate(feats)]


class PlainYOLO11Pose(nn.Module):
    def __init__(self, nc=1, kpt_shape=(5, 3), depth=0.50, width=0.25, max_channels=1024, c3k_all=False):
        super().__init__()
        ch = lambda c: _div8(min(c, max_channels) * width)  # noqa: E731
        rep = lambda n: max(round(n * depth), 1)  # noqa: E731
        c64, c128, c256, c512, c1024 = ch(64), ch(128), ch(256), ch(512), ch(1024)
        k = bool(c3k_all)
        self.b0, self.b1 = Conv(3, c64, 3, 2), Conv(c64, c128, 3, 2)
        self.b2 = C3k2(c128, c256, rep(2), k, 0.25)
        self.b3 = Conv(c256, c256, 3, 2)
        self.b4 = C3k2(c256, c512, rep(2), k, 0.25)
        self.b5 = Conv(c512, c512, 3, 2)
        self.b6 = C3k2(c512, c512, rep(2), True)
        self.b7 = Conv(c512, c1024, 3, 2)
        self.b8 = C3k2(c1024, c1024, rep(2), True)
        self.b9 = SPPF(c1024, c1024, 5)
        self.b10 = C2PSA(c1024, c1024, rep(2))
        self.up = nn.Upsample(scale_factor=2.0, mode="nearest")
        self.h13 = C3k2(c1024 + c512, c512, rep(2), k)
        self.h16 = C3k2(c512 + c512, c256, rep(2), k)
        self.h17 = Conv(c256, c256, 3, 2)
        self.h19 = C3k2(c256 + c512, c512, rep(2), k)
        self.h20 = Conv(c512, c512, 3, 2)
        self.h22 = C3k2(c512 + c1024, c1024, rep(2), True)
        self.head = PoseHead(nc, kpt_shape, (c256, c512, c1024))

    def forward(self, x):
        p3 = self.b4(self.b3(self.b2(self.b1(self.b0(x)))))
        p4 = self.b6(self.b5(p3))
        p5 = self.b10(self.b9(self.b8(self.b7(p4))))
        n4 = self.h13(torch.cat((self.up(p5), p4), 1))
        n3 = self.h16(torch.cat((self.up(n4), p3), 1))
        m4 = self.h19(torch.cat((self.h17(n3), n4), 1))
        m5 = self.h22(torch.cat((self.h20(m4), p5), 1))
        return self.head([n3, m4, m5])


def build_plain_yolo11n_pose(seed: int = 0, state_dict=None) -> PlainYOLO11Pose:
    """YOLO11n-pose with deterministic random weights (variance-preserving normal init; the class-logit bias is lowered so
    that a trained-detector-like fraction of anchors passes conf 0.5), or with the given state_dict."""
    model = PlainYOLO11Pose()
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    else:
        gen = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for m in model.modules():
                if isinstance(m, nn.Conv2d):
                    fan_in = m.in_channels // m.groups * m.kernel_size[0] * m.kernel_size[1]
                    m.weight.copy_(torch.randn(m.weight.shape, generator=gen) * (1.6 / math.sqrt(fan_in)))
                    m.bias.zero_()
        # give the raw head outputs trained-detector-like statistics (class logits ~ N(-6, 2): a fraction of a percent of the
        # anchors passes conf 0.5; DFL logits std 1.5; key-point offsets O(1)) on one synthetic WIDER-like image up-scaled 2x
        # like a SAHI slice — the same recipe and seed as the bench's GPU arm uses for its random-init network, so both arms
        # see the same detection load (the per-detection Python work of the reference path is a large part of its time)
        from fsd_b200.synthetic import make_image  # data generator only (numpy): no network code

        model.eval()
        with torch.no_grad():
            img, _ = make_image(10_000, 320, 320, seed=seed)
            sample = torch.from_numpy(img).permute(2, 0, 1)[None].float() / 255.0
            sample = torch.nn.functional.interpolate(sample, scale_factor=2.0, mode="bilinear", align_corners=False)
            for lvl, outs in enumerate(model(sample)):
                for branch, t, mean_to, std_to in ((model.head.cv2, outs[0], 1.0, 1.5), (model.head.cv3, outs[1], -6.0, 2.0),
                                                   (model.head.cv4, outs[2], 0.0, 1.0)):
                    conv = branch[lvl][-1]
                    scale = std_to / max(float(t.float().std()), 1e-6)
                    conv.weight.mul_(scale)
                    conv.bias.copy_((conv.bias - float(t.float().mean())) * scale + mean_to)
    for p in model.parameters():
        p.requires_grad_(False)
    return model.eval()
