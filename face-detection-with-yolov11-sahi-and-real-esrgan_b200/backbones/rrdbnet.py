"""RRDBNet (ESRGAN generator) in plain PyTorch, random-init.

Same constructor contract as [EXT basicsr==1.4.2] basicsr.archs.rrdbnet_arch.RRDBNet, which the reference builds at
utils/enhancer.py:99-129 (num_feat 64, num_block 23 or 6, num_grow_ch 32, scale 4/2/1); structure per SURVEY App. A.5.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


FAST_INFERENCE = True  # GPU fp16 inference: channels-last convolutions + the one-pass bias/LeakyReLU epilogue (fsd_bias_act_inplace)


def _fast(x):
    return FAST_INFERENCE and x.is_cuda and x.dtype == torch.float16 and not torch.is_grad_enabled()


def _conv_act(conv, x, act):
    """conv(x) + bias followed by LeakyReLU(0.2) (act="lrelu") or nothing (act="none").  On the fast path cuDNN produces the
    raw channels-last convolution and ONE hand-written pass applies bias and activation in place; eager PyTorch launches a
    bias-add kernel and an activation kernel after every convolution (2 x 115 convolutions in the 23-block network)."""
    if _fast(x) and conv.out_channels % 8 == 0:
        y = F.conv2d(x, conv.weight, None, conv.stride, conv.padding)
        if y.is_contiguous(memory_format=torch.channels_last):
            from ..ops import bias_act_

            return bias_act_(y, conv.bias, act, 0.2)
        y = y + conv.bias.view(1, -1, 1, 1)
    else:
        y = conv(x)
    return F.leaky_relu(y, 0.2) if act == "lrelu" else y


def pixel_unshuffle(x, scale):
    b, c, hh, hw = x.size()
    h, w = hh // scale, hw // scale
    return x.view(b, c, h, scale, w, scale).permute(0, 1, 3, 5, 2, 4).reshape(b, c * scale * scale, h, w)


class ResidualDenseBlock(nn.Module):
    def __init__(self, num_feat=64, num_grow_ch=32):
        super().__init__()
        self.conv1 = nn.Conv2d(num_feat, num_grow_ch, 3, 1, 1)
        self.conv2 = nn.Conv2d(num_feat + num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv3 = nn.Conv2d(num_feat + 2 * num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv4 = nn.Conv2d(num_feat + 3 * num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv5 = nn.Conv2d(num_feat + 4 * num_grow_ch, num_feat, 3, 1, 1)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)
        for m in (self.conv1, self.conv2, self.conv3, self.conv4, self.conv5):  # basicsr default_init_weights(0.1)
            nn.init.kaiming_normal_(m.weight)
            m.weight.data *= 0.1
            nn.init.zeros_(m.bias)

    def forward(self, x):
        x1 = _conv_act(self.conv1, x, "lrelu")
        x2 = _conv_act(self.conv2, torch.cat((x, x1), 1), "lrelu")
        x3 = _conv_act(self.conv3, torch.cat((x, x1, x2), 1), "lrelu")
        x4 = _conv_act(self.conv4, torch.cat((x, x1, x2, x3), 1), "lrelu")
        x5 = _conv_act(self.conv5, torch.cat((x, x1, x2, x3, x4), 1), "none")
        return torch.add(x, x5, alpha=0.2) if _fast(x) else x5 * 0.2 + x


class RRDB(nn.Module):
    def __init__(self, num_feat, num_grow_ch=32):
        super().__init__()
        self.rdb1 = ResidualDenseBlock(num_feat, num_grow_ch)
        self.rdb2 = ResidualDenseBlock(num_feat, num_grow_ch)
        self.rdb3 = ResidualDenseBlock(num_feat, num_grow_ch)

    def forward(self, x):
        y = self.rdb3(self.rdb2(self.rdb1(x)))
        return torch.add(x, y, alpha=0.2) if _fast(x) else y * 0.2 + x


class RRDBNet(nn.Module):
    def __init__(self, num_in_ch=3, num_out_ch=3, scale=4, num_feat=64, num_block=23, num_grow_ch=32):
        super().__init__()
        self.scale = scale
        if scale == 2:
            num_in_ch = num_in_ch * 4
        elif scale == 1:
            num_in_ch = num_in_ch * 16
        self.conv_first = nn.Conv2d(num_in_ch, num_feat, 3, 1, 1)
        self.body = nn.Sequential(*(RRDB(num_feat, num_grow_ch) for _ in range(num_block)))
        self.conv_body = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_up1 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_up2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_hr = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_last = nn.Conv2d(num_feat, num_out_ch, 3, 1, 1)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)

    def forward(self, x):
        if self.scale == 2:
            feat = pixel_unshuffle(x, 2)
        elif self.scale == 1:
            feat = pixel_unshuffle(x, 4)
        else:
            feat = x
        if _fast(feat):
            # channels-last from here on (the weights are converted once): cuDNN's tensor-core kernels are NHWC
            if not getattr(self, "_cl_weights", False):
                self.to(memory_format=torch.channels_last)
                self._cl_weights = True
            feat = feat.contiguous(memory_format=torch.channels_last)
        feat = _conv_act(self.conv_first, feat, "none")
        feat = feat + _conv_act(self.conv_body, self.body(feat), "none")
        feat = _conv_act(self.conv_up1, F.interpolate(feat, scale_factor=2, mode="nearest"), "lrelu")
        feat = _conv_act(self.conv_up2, F.interpolate(feat, scale_factor=2, mode="nearest"), "lrelu")
        out = self.conv_last(_conv_act(self.conv_hr, feat, "lrelu"))
        return out.contiguous() if _fast(x) else out
