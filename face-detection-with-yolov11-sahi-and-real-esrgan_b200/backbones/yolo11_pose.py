"""YOLO11-pose in plain PyTorch (random-init; there is no network for checkpoints).

Architecture follows ultralytics' `yolo11-pose.yaml` as recorded in SURVEY.md App. A.4.1 — the model the
reference loads through `ultralytics.YOLO(model_path)` at utils/yolo_wrapper.py:55 ([EXT ultralytics]).
Conv+BN pairs are built already fused (ultralytics fuses them before inference), so `Conv` = Conv2d(bias) + SiLU.

The module stops at the raw per-level head convolutions: `forward` returns
    [(box [B,64,h,w], cls [B,nc,h,w], kpt [B,15,h,w]) for stride in (8,16,32)]
and the decode (DFL, anchors, sigmoid, key-points, NMS) is Kernel 2/3's job (or oracle/yolo_head.py on the CPU).
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F


USE_S2D_STEM = not os.environ.get("FSD_NO_S2D_STEM")  # layers 0+1: fsd_stem_conv writes space-to-depth, layer 1 runs as a 2x2 stride-1 convolution (cuDNN fast path)
USE_POINTWISE_KERNEL = True  # fsd_pointwise_conv for the 1x1 layers it supports (False: cuDNN + fsd_bias_act everywhere)


def _make_divisible(x, divisor=8):
    return int(math.ceil(x / divisor) * divisor)


def _cat_buffer(x, channels, h=None, w=None):
    """Empty channels-last [N, channels, H, W] buffer whose channel slots the conv epilogues fill (no torch.cat pass)."""
    return torch.empty((x.shape[0], channels, x.shape[2] if h is None else h, x.shape[3] if w is None else w),
                       dtype=x.dtype, device=x.device, memory_format=torch.channels_last)


class Conv(nn.Module):
    def __init__(self, c1, c2, k=1, s=1, g=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2, groups=g, bias=True)
        self.act = nn.SiLU(inplace=True) if act else nn.Identity()

    def forward(self, x, out=None, residual=None, out2=None, up2=None):
        """act(conv(x) + bias) (+ residual); `up2` [N,C,2H,2W] additionally receives the result up-sampled 2x (nearest).  On the GPU in fp16 the convolution runs in cuDNN without bias and ONE
        hand-written pass (fsd_bias_act) applies bias, activation and the residual and stores the result into `out` —
        which may be a channel slot of a concat buffer — and the trailing channels into `out2` as well."""
        c = self.conv
        if (x.is_cuda and x.dtype == torch.float16 and c.out_channels % 8 == 0 and not torch.is_grad_enabled()):
            if (c.in_channels == 3 and c.out_channels == 16 and c.kernel_size == (3, 3) and c.stride == (2, 2)
                    and c.padding == (1, 1) and c.groups == 1 and isinstance(self.act, nn.SiLU) and out is None
                    and residual is None and out2 is None and up2 is None and x.shape[3] % 2 == 0
                    and x.is_contiguous(memory_format=torch.channels_last)):
                # the stem: cuDNN has no good kernel for a 3-channel channels-last input; one hand-written tensor-core
                # kernel does convolution + bias + SiLU (fsd_stem_conv)
                from ..ops import stem_conv

                cached = getattr(self, "_w_dense", None)  # (weight version, dense [o,c,ky,kx] copy of the channels-last weight)
                if cached is None or cached[0] != (c.weight._version, c.weight.data_ptr()):
                    cached = ((c.weight._version, c.weight.data_ptr()), c.weight.detach().contiguous().clone())
                    self._w_dense = cached
                return stem_conv(x, cached[1], c.bias)
            act = "silu" if isinstance(self.act, nn.SiLU) else "none"
            if (c.kernel_size == (1, 1) and c.stride == (1, 1) and c.groups == 1 and USE_POINTWISE_KERNEL and up2 is None
                    and x.stride(1) == 1 and x.shape[2] * x.shape[3] >= 1024):
                # low-intensity 1x1 layers: one kernel does GEMM + bias + activation (+ residual) into the slot
                from ..ops import pointwise_conv, pointwise_conv_supported, pointwise_tc_enabled, pointwise_tc_preferred

                # tensor-core kernel (tcgen05, k10_pointwise_tc.cu): every shape it takes.  The mma.sync kernel (FSD_K7_NO_TC=1) only
                # beats cuDNN + epilogue for K, N <= 64 (profiles/r1_kernels_k7.jsonl: 0.55 vs 0.34 of peak at 32->32; for N = 128
                # its SiLU epilogue is issue/MUFU-bound and the library pair is faster)
                tc = pointwise_tc_enabled() and pointwise_tc_preferred(c.in_channels, c.out_channels)
                if pointwise_conv_supported(c.in_channels, c.out_channels) and (tc or (c.in_channels <= 64 and c.out_channels <= 64)):
                    return pointwise_conv(x, c.weight, c.bias, act, out=out, residual=residual, out2=out2)
            if (c.kernel_size == (3, 3) and c.stride == (1, 1) and c.padding == (1, 1) and c.dilation == (1, 1) and c.groups == 1
                    and out2 is None and up2 is None and x.stride(1) == 1 and x.shape[2] * x.shape[3] >= 256):
                # dense 3x3 layers: implicit GEMM on the tensor cores with the epilogue fused (fsd_conv3x3)
                from ..ops import conv3x3, conv3x3_preferred, conv3x3_tap_major, conv3x3_tc_enabled

                if conv3x3_tc_enabled() and conv3x3_preferred(c.in_channels, c.out_channels):
                    cached = getattr(self, "_w_taps", None)  # (weight version, tap-major copy)
                    if cached is None or cached[0] != (c.weight._version, c.weight.data_ptr()):
                        cached = ((c.weight._version, c.weight.data_ptr()), conv3x3_tap_major(c.weight))
                        self._w_taps = cached
                    return conv3x3(x, cached[1], c.bias, act, out=out, residual=residual)
            if (c.kernel_size == (3, 3) and c.stride == (1, 1) and c.padding == (1, 1) and c.dilation == (1, 1)
                    and c.groups == c.in_channels == c.out_channels and residual is None and out2 is None and up2 is None
                    and x.stride(1) == 1 and not os.environ.get("FSD_NO_DWCONV")):
                # depth-wise 3x3 layers: one stencil kernel with the epilogue fused (fsd_dwconv3x3)
                from ..ops import dwconv3x3, dwconv3x3_tap_major

                cached = getattr(self, "_w_taps", None)
                if cached is None or cached[0] != (c.weight._version, c.weight.data_ptr()):
                    cached = ((c.weight._version, c.weight.data_ptr()), dwconv3x3_tap_major(c.weight))
                    self._w_taps = cached
                return dwconv3x3(x, cached[1], c.bias, act, out=out)
            y = F.conv2d(x, c.weight, None, c.stride, c.padding, c.dilation, c.groups)
            if y.is_contiguous(memory_format=torch.channels_last):
                from ..ops import bias_act

                return bias_act(y, c.bias, act, out=out, residual=residual, out2=out2, up2=up2)
            y = self.act(y + c.bias.view(1, -1, 1, 1))
        else:
            y = self.act(c(x))
        if residual is not None:
            y = residual + y
        if out2 is not None:
            out2.copy_(y[:, y.shape[1] - out2.shape[1]:])
        if up2 is not None:
            up2.copy_(F.interpolate(y, scale_factor=2.0, mode="nearest"))
        if out is not None:
            out.copy_(y)
            return out
        return y


class DWConv(Conv):
    def __init__(self, c1, c2, k=1, s=1, act=True):
        super().__init__(c1, c2, k, s, g=math.gcd(c1, c2), act=act)


class Bottleneck(nn.Module):
    def __init__(self, c1, c2, shortcut=True, k=(3, 3), e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, k[0], 1)
        self.cv2 = Conv(c_, c2, k[1], 1)
        self.add = shortcut and c1 == c2

    def forward(self, x, out=None):
        c1, c2 = self.cv1.conv, self.cv2.conv
        if (c1.out_channels == 8 and x.is_cuda and x.dtype == torch.float16 and not torch.is_grad_enabled() and x.stride(1) == 1
                and c1.kernel_size == (3, 3) and c2.kernel_size == (3, 3) and c1.stride == (1, 1) and c2.stride == (1, 1)
                and c1.groups == 1 and c2.groups == 1 and isinstance(self.cv1.act, nn.SiLU) and isinstance(self.cv2.act, nn.SiLU)
                and not os.environ.get("FSD_NO_CONV3_TC")):
            # An 8-channel hidden tensor (yolo11n: C3k2 of layer 2) is below the tensor-core kernel's K = 16: carry it as 16 channels
            # whose upper half is exactly zero (zero weight rows and bias -> SiLU(0) = 0) and give cv2 zero weight columns there —
            # the same sums, both convolutions on fsd_conv3x3 instead of cuDNN + epilogue for cv2.
            from ..ops import conv3x3, conv3x3_supported, conv3x3_tap_major

            if conv3x3_supported(c1.in_channels, 16) and conv3x3_supported(16, c2.out_channels) and c1.in_channels <= 64:
                key = (c1.weight._version, c1.weight.data_ptr(), c2.weight._version, c2.weight.data_ptr())
                cached = getattr(self, "_padded", None)
                if cached is None or cached[0] != key:
                    w1 = torch.cat([c1.weight.detach(), torch.zeros_like(c1.weight)], dim=0)           # [16, K, 3, 3]
                    b1 = torch.cat([c1.bias.detach(), torch.zeros_like(c1.bias)], dim=0)
                    w2 = torch.cat([c2.weight.detach(), torch.zeros_like(c2.weight)], dim=1)           # [N, 16, 3, 3]
                    cached = (key, conv3x3_tap_major(w1), b1.contiguous(), conv3x3_tap_major(w2))
                    self._padded = cached
                hidden = conv3x3(x, cached[1], cached[2], "silu")
                return conv3x3(hidden, cached[3], c2.bias, "silu", out=out, residual=x if self.add else None)
        return self.cv2(self.cv1(x), out=out, residual=x if self.add else None)


class C3k(nn.Module):
    def __init__(self, c1, c2, n=2, shortcut=True, e=0.5, k=3):
        super().__init__()
        c_ = int(c2 * e)
        self.c_ = c_
        self.cv1, self.cv2, self.cv3 = Conv(c1, c_, 1, 1), Conv(c1, c_, 1, 1), Conv(2 * c_, c2, 1)
        self.m = nn.ModuleList(Bottleneck(c_, c_, shortcut, k=(k, k), e=1.0) for _ in range(n))

    def forward(self, x, out=None):
        buf = _cat_buffer(x, 2 * self.c_)
        y = self.cv1(x)
        for i, m in enumerate(self.m):
            y = m(y, out=buf[:, :self.c_] if i == len(self.m) - 1 else None)
        if len(self.m) == 0:
            buf[:, :self.c_].copy_(y)
        self.cv2(x, out=buf[:, self.c_:])
        return self.cv3(buf, out=out)


class C3k2(nn.Module):
    def __init__(self, c1, c2, n=1, c3k=False, e=0.5, shortcut=True):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(C3k(self.c, self.c, 2, shortcut) if c3k else Bottleneck(self.c, self.c, shortcut)
                               for _ in range(n))

    def forward(self, x, out=None, out2=None, up2=None):
        c, n = self.c, len(self.m)
        buf = _cat_buffer(x, (2 + n) * c)
        y = torch.empty((x.shape[0], c, x.shape[2], x.shape[3]), dtype=x.dtype, device=x.device,
                        memory_format=torch.channels_last)
        # cv1 fills slots 0-1 of the concat buffer; its second half is also stored densely as the input of m[0]
        self.cv1(x, out=buf[:, :2 * c], out2=y)
        for i, m in enumerate(self.m):
            y = m(y, out=buf[:, (2 + i) * c:(3 + i) * c])
            if i + 1 < n:
                y = y.contiguous(memory_format=torch.channels_last)
        return self.cv2(buf, out=out, out2=out2, up2=up2)


class SPPF(nn.Module):
    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.c_ = c_
        self.cv1, self.cv2 = Conv(c1, c_, 1, 1), Conv(c_ * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)

    def forward(self, x):
        buf = _cat_buffer(x, 4 * self.c_)
        h, w = x.shape[2:]
        if (x.is_cuda and x.dtype == torch.float16 and self.c_ % 8 == 0 and self.m.kernel_size == 5
                and 2 * h * w * 16 <= 200 * 1024 and not torch.is_grad_enabled()):
            from ..ops import sppf_pool_

            self.cv1(x, out=buf[:, :self.c_])
            sppf_pool_(buf)  # slots 1..3 = m(y), m(m(y)), m(m(m(y))) in one launch
        else:
            y = self.cv1(x)
            buf[:, :self.c_].copy_(y)
            for i in range(1, 4):
                y = self.m(y)
                buf[:, i * self.c_:(i + 1) * self.c_].copy_(y)
        return self.cv2(buf)


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, attn_ratio=0.5):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.key_dim = int(self.head_dim * attn_ratio)
        self.scale = self.key_dim ** -0.5
        h = dim + self.key_dim * num_heads * 2
        self.qkv = Conv(dim, h, 1, act=False)
        self.proj = Conv(dim, dim, 1, act=False)
        self.pe = Conv(dim, dim, 3, 1, g=dim, act=False)

    def forward(self, x, residual=None):
        """ultralytics Attention.forward: softmax(q^T k * scale) applied to v, plus the depth-wise positional conv of v.
        Evaluated with torch's fused scaled_dot_product_attention on the channels-last memory of the qkv convolution
        ([B, N, heads, 2*key_dim + head_dim] is a free view of it), instead of matmul -> scale -> softmax -> matmul
        (three [B, heads, N, N] round trips through HBM)."""
        B, C, H, W = x.shape
        N = H * W
        qkv = self.qkv(x)  # [B, heads * (2 kd + hd), H, W]; channel = head * (2 kd + hd) + d
        t = qkv.permute(0, 2, 3, 1).reshape(B, N, self.num_heads, 2 * self.key_dim + self.head_dim)
        q, k, v = t.split([self.key_dim, self.key_dim, self.head_dim], dim=3)  # [B, N, heads, d], unit last stride
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), scale=self.scale)
        o = o.transpose(1, 2).reshape(B, H, W, C).permute(0, 3, 1, 2)    # [B, C, H, W], channel = head * hd + d
        vv = v.reshape(B, H, W, C).permute(0, 3, 1, 2)                    # same channel order, channels-last dense
        if not o.is_contiguous(memory_format=torch.channels_last):
            o = o.contiguous(memory_format=torch.channels_last)
        x = self.pe(vv, residual=o)
        return self.proj(x, residual=residual)


class PSABlock(nn.Module):
    def __init__(self, c, attn_ratio=0.5, num_heads=4):
        super().__init__()
        self.attn = Attention(c, num_heads=num_heads, attn_ratio=attn_ratio)
        self.ffn = nn.Sequential(Conv(c, c * 2, 1), Conv(c * 2, c, 1, act=False))

    def forward(self, x, out=None):
        x = self.attn(x, residual=x)
        return self.ffn[1](self.ffn[0](x), out=out, residual=x)


class C2PSA(nn.Module):
    def __init__(self, c1, c2, n=1, e=0.5):
        super().__init__()
        assert c1 == c2
        self.c = int(c1 * e)
        self.cv1, self.cv2 = Conv(c1, 2 * self.c, 1, 1), Conv(2 * self.c, c1, 1)
        self.m = nn.ModuleList(PSABlock(self.c, attn_ratio=0.5, num_heads=max(1, self.c // 64)) for _ in range(n))

    def forward(self, x, out=None, out2=None, up2=None):
        c = self.c
        buf = _cat_buffer(x, 2 * c)
        b = torch.empty((x.shape[0], c, x.shape[2], x.shape[3]), dtype=x.dtype, device=x.device,
                        memory_format=torch.channels_last)
        self.cv1(x, out=buf, out2=b)  # slot 0 = a, slot 1 is overwritten by m(b) below
        for i, m in enumerate(self.m):
            b = m(b, out=buf[:, c:] if i == len(self.m) - 1 else None)
        return self.cv2(buf, out=out, out2=out2, up2=up2)


def _up_cat(a, b):
    """cat(nearest-2x(a), b): one fused pass on the GPU (fsd_upsample2x_concat), plain torch elsewhere."""
    if (a.is_cuda and a.dtype == torch.float16 and a.shape[1] % 8 == 0 and b.shape[1] % 8 == 0
            and a.is_contiguous(memory_format=torch.channels_last) and b.is_contiguous(memory_format=torch.channels_last)
            and not torch.is_grad_enabled()):
        from ..ops import upsample2x_concat

        return upsample2x_concat(a, b)
    return torch.cat((F.interpolate(a, scale_factor=2.0, mode="nearest"), b), 1)


class PoseHead(nn.Module):
    """ultralytics Pose(Detect) head up to (and including) the last 1x1 convolutions of cv2/cv3/cv4."""

    def __init__(self, nc, kpt_shape, ch):
        super().__init__()
        self.nc, self.kpt_shape, self.reg_max = nc, tuple(kpt_shape), 16
        self.nk = kpt_shape[0] * kpt_shape[1]
        c2, c3 = max(16, ch[0] // 4, self.reg_max * 4), max(ch[0], min(nc, 100))
        c4 = max(ch[0] // 4, self.nk)
        self.cv2 = nn.ModuleList(nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * self.reg_max, 1))
                                 for x in ch)
        self.cv3 = nn.ModuleList(nn.Sequential(nn.Sequential(DWConv(x, x, 3), Conv(x, c3, 1)),
                                               nn.Sequential(DWConv(c3, c3, 3), Conv(c3, c3, 1)),
                                               nn.Conv2d(c3, nc, 1)) for x in ch)
        self.cv4 = nn.ModuleList(nn.Sequential(Conv(x, c4, 3), Conv(c4, c4, 3), nn.Conv2d(c4, self.nk, 1)) for x in ch)

    @staticmethod
    def _branch(seq, x):
        """Sequential(..., nn.Conv2d 1x1 with bias): the last convolution's bias goes through the one-pass epilogue when
        its channel count allows (box: 64), instead of torch's separate broadcast add."""
        last = seq[-1]
        for m in seq[:-1]:
            x = m(x)
        if (x.is_cuda and x.dtype == torch.float16 and last.out_channels % 8 == 0 and not torch.is_grad_enabled()):
            if USE_POINTWISE_KERNEL and x.stride(1) == 1 and x.shape[2] * x.shape[3] >= 1024:
                from ..ops import pointwise_conv, pointwise_conv_supported

                if pointwise_conv_supported(last.in_channels, last.out_channels) and last.in_channels <= 64 and last.out_channels <= 64:
                    return pointwise_conv(x, last.weight, last.bias, "none")
            y = F.conv2d(x, last.weight, None, last.stride, last.padding, last.dilation, last.groups)
            if y.is_contiguous(memory_format=torch.channels_last):
                from ..ops import bias_act

                return bias_act(y, last.bias, "none")
            return y + last.bias.view(1, -1, 1, 1)
        return last(x)

    def forward(self, feats):
        return [(self._branch(self.cv2[i], x), self._branch(self.cv3[i], x), self._branch(self.cv4[i], x))
                for i, x in enumerate(feats)]


class YOLO11Pose(nn.Module):
    strides = (8, 16, 32)

    def __init__(self, nc=1, kpt_shape=(5, 3), depth=0.50, width=0.25, max_channels=1024, c3k_all=False):
        """yolo11.yaml scales: n (0.50, 0.25, 1024), s (0.50, 0.50, 1024), m (0.50, 1.00, 512), l (1.00, 1.00, 512),
        x (1.00, 1.50, 512); the m/l/x scales use C3k inside every C3k2 (`c3k_all`, ultralytics parse_model)."""
        super().__init__()
        self.arch = dict(nc=nc, kpt_shape=tuple(kpt_shape), depth=depth, width=width, max_channels=max_channels, c3k_all=c3k_all)
        k = bool(c3k_all)
        ch = lambda c: _make_divisible(min(c, max_channels) * width, 8)  # noqa: E731
        rep = lambda n: max(round(n * depth), 1)  # noqa: E731
        c64, c128, c256, c512, c1024 = ch(64), ch(128), ch(256), ch(512), ch(1024)
        self.nc, self.kpt_shape = nc, tuple(kpt_shape)
        # backbone
        self.b0 = Conv(3, c64, 3, 2)
        self.b1 = Conv(c64, c128, 3, 2)
        self.b2 = C3k2(c128, c256, rep(2), k, 0.25)
        self.b3 = Conv(c256, c256, 3, 2)
        self.b4 = C3k2(c256, c512, rep(2), k, 0.25)
        self.b5 = Conv(c512, c512, 3, 2)
        self.b6 = C3k2(c512, c512, rep(2), True)
        self.b7 = Conv(c512, c1024, 3, 2)
        self.b8 = C3k2(c1024, c1024, rep(2), True)
        self.b9 = SPPF(c1024, c1024, 5)
        self.b10 = C2PSA(c1024, c1024, rep(2))
        # neck
        self.h13 = C3k2(c1024 + c512, c512, rep(2), k)
        self.h16 = C3k2(c512 + c512, c256, rep(2), k)
        self.h17 = Conv(c256, c256, 3, 2)
        self.h19 = C3k2(c256 + c512, c512, rep(2), k)
        self.h20 = Conv(c512, c512, 3, 2)
        self.h22 = C3k2(c512 + c1024, c1024, rep(2), True)
        self.head = PoseHead(nc, kpt_shape, (c256, c512, c1024))

    def _stem(self, x):
        """b1(b0(x)).  On the GPU in fp16: layer 0 is the hand-written stem kernel, which can store its output already
        folded space-to-depth ([E, 64, H/4 + 1, W/4 + 1], zero first row / column); on that tensor layer 1
        (Conv(16, 32, 3, 2)) is algebraically a 2x2 stride-1 convolution over 64 channels — same products, same sums — for
        which cuDNN has a tensor-core kernel (0.33 ms vs 0.52 ms per 96 inputs, benchmarks/b1_s2d_probe.py)."""
        c0, c1 = self.b0.conv, self.b1.conv
        if (USE_S2D_STEM and x.is_cuda and x.dtype == torch.float16 and not torch.is_grad_enabled()
                and x.is_contiguous(memory_format=torch.channels_last) and x.shape[2] % 4 == 0 and x.shape[3] % 4 == 0
                and (c0.in_channels, c0.out_channels, c0.kernel_size, c0.stride, c0.padding, c0.groups) == (3, 16, (3, 3), (2, 2), (1, 1), 1)
                and (c1.in_channels, c1.kernel_size, c1.stride, c1.padding, c1.groups) == (16, (3, 3), (2, 2), (1, 1), 1)
                and c1.out_channels % 8 == 0 and isinstance(self.b0.act, nn.SiLU) and isinstance(self.b1.act, nn.SiLU)):
            from ..ops import bias_act, stem_conv

            key = (c0.weight._version, c0.weight.data_ptr(), c1.weight._version, c1.weight.data_ptr())
            cached = getattr(self, "_stem_weights", None)
            if cached is None or cached[0] != key:
                w0 = c0.weight.detach().contiguous().clone()
                w1 = c1.weight.detach()
                w2 = torch.zeros((c1.out_channels, 64, 2, 2), dtype=w1.dtype, device=w1.device)
                tap = {0: (0, 1), 1: (1, 0), 2: (1, 1)}  # kernel row/col k -> (2x2 tap, parity inside the 2x2 block)
                for ky in range(3):
                    a, dy = tap[ky]
                    for kx in range(3):
                        b, dx = tap[kx]
                        w2[:, (dy * 2 + dx) * 16:(dy * 2 + dx + 1) * 16, a, b] = w1[:, :, ky, kx]
                cached = (key, w0, w2.contiguous(memory_format=torch.channels_last), w2.permute(2, 3, 0, 1).contiguous())
                self._stem_weights = cached
            s = stem_conv(x, cached[1], c0.bias, space_to_depth=True)
            from ..ops import conv2x2, conv2x2_supported, conv3x3_tc_enabled

            if conv3x3_tc_enabled() and conv2x2_supported(64, c1.out_channels):
                return conv2x2(s, cached[3], c1.bias, "silu")  # tensor-core implicit GEMM with the epilogue fused
            y = F.conv2d(s, cached[2], None, 1, 0)
            if y.is_contiguous(memory_format=torch.channels_last):
                return bias_act(y, c1.bias, "silu")
            return self.b1.act(y + c1.bias.view(1, -1, 1, 1))
        return self.b1(self.b0(x))

    def forward(self, x):
        """Every concatenation of the FPN / PAN neck is a pre-allocated buffer that its producers fill: the conv epilogues
        store p3 / p4 (second destination), the nearest-2x up-sampled p5 / n4 (`up2`) and the PAN inputs straight into their
        channel slots, so neither `torch.cat` nor an up-sampling kernel runs."""
        x = self._stem(x)
        down = lambda v: (v - 1) // 2 + 1  # noqa: E731  (3x3 stride-2 pad-1 convolution)
        h8, w8 = down(x.shape[2]), down(x.shape[3])
        h16, w16 = down(h8), down(w8)
        h32, w32 = down(h16), down(w16)
        if (2 * h32, 2 * w32, 2 * h16, 2 * w16) != (h16, w16, h8, w8):
            raise ValueError("the network input must be a multiple of 32 (ultralytics pads to the stride before inference)")
        c3, c4, c5 = self.b4.cv2.conv.out_channels, self.b6.cv2.conv.out_channels, self.b10.cv2.conv.out_channels
        cn4 = self.h13.cv2.conv.out_channels
        c17, c20 = self.h17.conv.out_channels, self.h20.conv.out_channels
        cat16 = _cat_buffer(x, cn4 + c3, h8, w8)      # cat(up(n4), p3)
        p3 = self.b4(self.b3(self.b2(x)), out2=cat16[:, cn4:])
        cat13 = _cat_buffer(x, c5 + c4, h16, w16)     # cat(up(p5), p4)
        p4 = self.b6(self.b5(p3), out2=cat13[:, c5:])
        cat22 = _cat_buffer(x, c20 + c5, h32, w32)    # cat(h20(m4), p5)
        self.b10(self.b9(self.b8(self.b7(p4))), out=cat22[:, c20:], up2=cat13[:, :c5])      # p5
        cat19 = _cat_buffer(x, c17 + cn4, h16, w16)   # cat(h17(n3), n4)
        self.h13(cat13, out=cat19[:, c17:], up2=cat16[:, :cn4])                               # n4
        n3 = self.h16(cat16)
        self.h17(n3, out=cat19[:, :c17])
        m4 = self.h19(cat19)
        self.h20(m4, out=cat22[:, :c20])
        m5 = self.h22(cat22)
        return self.head([n3, m4, m5])

    @staticmethod
    def anchors_for(h, w):
        """number of anchors for a network input of h x w"""
        return sum((h // s) * (w // s) for s in YOLO11Pose.strides)


def _init_variance_preserving(model: nn.Module, gen: torch.Generator):
    """Random init that keeps activations O(1) through ~100 SiLU convs so fp16 inference does not flush to zero."""
    for m in model.modules():
        if isinstance(m, nn.Conv2d):
            fan_in = m.in_channels // m.groups * m.kernel_size[0] * m.kernel_size[1]
            std = 1.6 / math.sqrt(fan_in)
            with torch.no_grad():
                m.weight.copy_(torch.randn(m.weight.shape, generator=gen) * std)
                m.bias.zero_()


@torch.no_grad()
def calibrate_head(model: YOLO11Pose, sample: torch.Tensor, cls_logit_mean=-6.0, cls_logit_std=2.0,
                   box_logit_std=1.5, kpt_std=1.0):
    """Rescale the last 1x1 head convolutions so that random-init outputs look like a trained detector's:
    a fraction of a percent of anchors pass conf=0.5, DFL bins are peaked, key-points are O(1) offsets.
    Deterministic given the weights and `sample` (a fixed seeded image batch)."""
    model.eval()
    outs = model(sample)
    for lvl, (box, cls, kpt) in enumerate(outs):
        for branch, t, tgt_mean, tgt_std in ((model.head.cv2, box, 1.0, box_logit_std),
                                             (model.head.cv3, cls, cls_logit_mean, cls_logit_std),
                                             (model.head.cv4, kpt, 0.0, kpt_std)):
            conv = branch[lvl][-1]
            std = float(t.float().std())
            mean = float(t.float().mean())
            scale = tgt_std / max(std, 1e-6)
            conv.weight.mul_(scale)
            conv.bias.copy_((conv.bias - mean) * scale + tgt_mean)
    return model


def build_yolo11n_pose(seed: int = 0, calibrate: bool = True, nc: int = 1, kpt_shape=(5, 3)) -> YOLO11Pose:
    """YOLO11n-pose (depth 0.50, width 0.25), nc=1 'face', 5 key-points x (x, y, conf); deterministic random init."""
    gen = torch.Generator().manual_seed(seed)
    model = YOLO11Pose(nc=nc, kpt_shape=kpt_shape)
    _init_variance_preserving(model, gen)
    if calibrate:
        # calibrate on one synthetic WIDER-like image, up-scaled 2x like a SAHI slice at imgsz 1024
        from ..synthetic import make_image

        img, _ = make_image(10_000, 320, 320, seed=seed)
        sample = torch.from_numpy(img).permute(2, 0, 1)[None].float() / 255.0
        sample = torch.nn.functional.interpolate(sample, scale_factor=2.0, mode="bilinear", align_corners=False)
        calibrate_head(model, sample)
    for p in model.parameters():
        p.requires_grad_(False)
    return model.eval()
