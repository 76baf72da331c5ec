"""MobileNetV2 backbone with the reference's constructor, attribute names and state_dict keys
(modeling/backbone/mobilenet.py:71-145), executed on the fused sm_100a kernels.

The nn.Conv2d / BatchNorm / ReLU6 children are parameter containers laid out exactly like the
reference (same creation order, so the same torch seed yields the same initial weights); the
forward pass never calls them, it walks them with engine.InvertedResidual.
"""
import os

import torch
import torch.nn as nn

from .. import sync_batchnorm  # noqa: F401
from ... import _lib as L
from ...engine import ConvBNAct, InvertedResidual as BlockRun, bn_backward
from ...runtime import RunBase, call_module, init_reference_weights

# expansion t, channels c, repeats n, stride s  (mobilenet.py:78-87 of the reference)
_BLOCK_TABLE = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2),
                (6, 320, 1, 1))


def conv_bn(inp, oup, stride, BatchNorm):
    return nn.Sequential(nn.Conv2d(inp, oup, 3, stride, 1, bias=False), BatchNorm(oup), nn.ReLU6(inplace=True))


class InvertedResidual(nn.Module):
    """Parameter container of one block; `conv` has the reference's Sequential indices."""

    def __init__(self, inp, oup, stride, dilation, expand_ratio, BatchNorm):
        super().__init__()
        assert stride in [1, 2]
        self.stride = stride
        self.kernel_size = 3
        self.dilation = dilation
        hidden = round(inp * expand_ratio)
        self.use_res_connect = stride == 1 and inp == oup
        layers = []
        if expand_ratio != 1:
            layers += [nn.Conv2d(inp, hidden, 1, 1, 0, 1, bias=False), BatchNorm(hidden), nn.ReLU6(inplace=True)]
        layers += [nn.Conv2d(hidden, hidden, 3, stride, 0, dilation, groups=hidden, bias=False),
                   BatchNorm(hidden), nn.ReLU6(inplace=True),
                   nn.Conv2d(hidden, oup, 1, 1, 0, 1, bias=False), BatchNorm(oup)]
        self.conv = nn.Sequential(*layers)

    def forward(self, x):
        raise L.S2RError("InvertedResidual is executed by its MobileNetV2 parent on the fused kernels")


class MobileNetV2Run(RunBase):
    """One forward/backward of the backbone on Act views."""
    raw_inputs = True   # the stem reads the NCHW fp32 image through the patch kernel

    def __init__(self, mod):
        feats = mod.features
        self.stem = ConvBNAct(feats[0][0], feats[0][1], L.ACT_RELU6)
        self.blocks = [BlockRun(m) for m in list(feats)[1:]]
        self.n_low = 3  # features[1:4] produce the low-level feature (mobilenet.py:116)

    def forward(self, cx, x):
        z0, st0 = self.stem.forward_raw(cx, x)
        h = cx.tr('block1', self.blocks[0].forward(cx, z0, lazy=st0))
        low = None
        for i, b in enumerate(self.blocks[1:], start=1):
            h = cx.tr('block%d' % (i + 1), b.forward(cx, h))
            if i + 1 == self.n_low:
                low = h
        self.z0_st0 = (z0, st0)
        return h, low

    def backward(self, cx, douts, need=None):
        dhigh, dlow = douts
        d = dhigh
        nb = len(self.blocks)
        for i in range(nb - 1, 0, -1):
            if i + 1 == self.n_low and dlow is not None:
                if d is None:
                    d = dlow
                else:
                    L.call("s2r_add_bf16", d.vp(), dlow.vp(), d.P * d.pitch, cx.stream)
            if d is None:
                continue
            d, _ = self.blocks[i].backward(cx, d)
        if d is None:
            return None
        g0, bsums = self.blocks[0].backward(cx, d)
        z0, st0 = self.z0_st0
        self.z0_st0 = None
        if st0.frozen and bsums is None:
            bsums = cx.f64(2 * z0.C)
        dz0 = cx.new(z0.N, z0.H, z0.W, z0.C)
        bn_backward(cx, self.stem.bn, g0, z0, st0, L.ACT_NONE, dz0, presummed=bsums)
        self.stem.backward_raw(cx, dz0, need_dx=False)
        return None


class MobileNetV2(nn.Module):
    def __init__(self, output_stride=8, BatchNorm=None, width_mult=1., pretrained=True):
        super().__init__()
        input_channel = int(32 * width_mult)
        current_stride = 2
        rate = 1
        feats = [conv_bn(3, input_channel, 2, BatchNorm)]
        for t, c, n, s in _BLOCK_TABLE:
            # once the requested output stride is reached, strides turn into dilation (mobilenet.py:95-102)
            if current_stride == output_stride:
                stride, dilation = 1, rate
                rate *= s
            else:
                stride, dilation = s, 1
                current_stride *= s
            oup = int(c * width_mult)
            for i in range(n):
                feats.append(InvertedResidual(input_channel, oup, stride if i == 0 else 1, dilation, t, BatchNorm))
                input_channel = oup
        self.features = nn.Sequential(*feats)
        init_reference_weights(self)
        if pretrained:
            self._load_pretrained_model()
        self.low_level_features = self.features[0:4]
        self.high_level_features = self.features[4:]
        self._s2r_has_sync_bn = bool(getattr(BatchNorm, "_s2r_sync", False))

    def forward(self, x):
        return call_module(self, lambda: MobileNetV2Run(self), (x,))

    def _load_pretrained_model(self):
        # mobilenet.py:124-132: the checkpoint is optional here (the reference tree ships without it)
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'mobilenet_VOC.pth')
        if not os.path.isfile(path):
            return
        pretrain_dict = torch.load(path, map_location='cpu')
        state_dict = self.state_dict()
        state_dict.update({k: v for k, v in pretrain_dict.items() if k in state_dict})
        self.load_state_dict(state_dict)
