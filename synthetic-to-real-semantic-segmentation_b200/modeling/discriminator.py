"""FCDiscriminator with the reference's names and state_dict keys (modeling/discriminator.py:6-35):
five biased 4x4 stride-2 convolutions, LeakyReLU(0.2) between them, as tap-GEMMs over the four
input-parity views with bias + LeakyReLU fused in the epilogue; the backward fuses the LeakyReLU
mask into the next layer's data-gradient epilogue.
"""
import torch.nn as nn

from .. import _lib as L
from ..engine import conv_fwd, conv_dgrad, conv_wgrad, bias_grad, conv_out_hw, round_up, im2col, patch_weight, Act
from ..runtime import RunBase, call_module


class FCDiscriminatorRun(RunBase):
    raw_inputs = True   # conv1 reads the NCHW fp32 input through the patch kernel (19 channels, 16 taps)

    def __init__(self, mod):
        self.convs = [mod.conv1, mod.conv2, mod.conv3, mod.conv4, mod.classifier]
        self.slope = mod.leaky_relu.negative_slope

    def forward(self, cx, x):
        self.in_shape = (x.N, x.H, x.W, x.C)
        acts = [im2col(cx, x, 4, 4, 2, 1)]   # conv1 as a pointwise GEMM over 4x4 patches
        h = x
        for i, c in enumerate(self.convs):
            last = i == len(self.convs) - 1
            Cout = c.weight.shape[0]
            OH, OW = conv_out_hw(h.H, h.W, 4, 4, 2, 1, 1)
            y = cx.new(h.N, OH, OW, round_up(Cout, 8), zero=Cout % 8 != 0)
            y.C = Cout
            if i == 0:
                conv_fwd(cx, acts[0], patch_weight(c.weight), y, bias=c.bias, act=L.ACT_LEAKY, slope=self.slope)
            else:
                conv_fwd(cx, h, c.weight, y, stride=2, pad=1, bias=c.bias,
                         act=L.ACT_NONE if last else L.ACT_LEAKY, slope=self.slope)
            acts.append(y)
            h = y
        self.acts = acts
        return h

    def backward(self, cx, douts, need=None):
        d = douts[0] if isinstance(douts, tuple) else douts
        acts = self.acts
        self.acts = None
        need_dx = True if need is None else bool(need[0])
        for i in range(len(self.convs) - 1, -1, -1):
            c = self.convs[i]
            xin = acts[i]
            if c.weight.requires_grad:
                if i == 0:
                    conv_wgrad(cx, xin, d, patch_weight(c.weight), grad_param=c.weight)
                else:
                    conv_wgrad(cx, xin, d, c.weight, stride=2, pad=1)
            if c.bias is not None and c.bias.requires_grad:
                bias_grad(cx, d, c.bias)
            if i == 0 and not need_dx:
                return None
            if i == 0:
                N, H, W, Cc = self.in_shape   # gradient w.r.t. the input pixels: the ordinary 4x4 data gradient
                dx = cx.new(N, H, W, round_up(Cc, 8))
                dx.C = Cc
            else:
                dx = cx.new(xin.N, xin.H, xin.W, xin.pitch)
                dx.C = xin.C
            # dz_{i-1} = dgrad * leaky'(y_{i-1}); the first layer's input has no activation
            conv_dgrad(cx, d, c.weight, dx, stride=2, pad=1, aux=xin if i > 0 else None,
                       aux_mode=L.AUX_LEAKY_MASK if i > 0 else L.AUX_NONE, slope=self.slope)
            d = dx
        return d


class FCDiscriminator(nn.Module):
    def __init__(self, num_classes, ndf=64):
        super().__init__()
        self.conv1 = nn.Conv2d(num_classes, ndf, kernel_size=4, stride=2, padding=1)
        self.conv2 = nn.Conv2d(ndf, ndf * 2, kernel_size=4, stride=2, padding=1)
        self.conv3 = nn.Conv2d(ndf * 2, ndf * 4, kernel_size=4, stride=2, padding=1)
        self.conv4 = nn.Conv2d(ndf * 4, ndf * 8, kernel_size=4, stride=2, padding=1)
        self.classifier = nn.Conv2d(ndf * 8, 1, kernel_size=4, stride=2, padding=1)
        self.leaky_relu = nn.LeakyReLU(negative_slope=0.2, inplace=True)

    def forward(self, x):
        return call_module(self, lambda: FCDiscriminatorRun(self), (x,))
