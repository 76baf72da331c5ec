"""DeepLabV3+ decoder with the reference's names and state_dict keys (modeling/decoder.py:7-57).

low-level 1x1 (24->48)+BN+ReLU and the x4 bilinear up-sampling of the ASPP output write into the
two channel slices of one NHWC concat buffer; then 3x3 304->256, 3x3 256->256 (each BN, ReLU,
Dropout) and the biased 1x1 classifier.
"""
import torch
import torch.nn as nn

from .. import _lib as L
from ..engine import ConvBNAct, conv_fwd, conv_dgrad, conv_wgrad, bias_grad, round_up, _vp
from ..functional import pop_pending_scale
from ..runtime import RunBase, call_module, init_reference_weights


class UpsampledLogits(object):
    """Module-boundary conversion for runs whose output is F.interpolate(logits, size, mode='bilinear',
    align_corners=True) of the decoder's NHWC bf16 logits (deeplab.py:31; train.py:184,194,270): the up-sampling is
    fused with the NHWC bf16 -> NCHW fp32 export, its gradient with the import (and with the deferred mean-reduction
    factor of the cross entropy, functional.DEFER_CE_SCALE).  Set self.out_hw before export()."""

    def export(self, cx, i, a):
        H, W = self.out_hw
        self.small = (a.N, a.H, a.W, a.C, a.pitch)
        y = torch.empty((a.N, a.C, H, W), dtype=torch.float32, device=cx.device)
        L.call("s2r_upsample_bilinear_nhwc_to_nchw", a.vp(), a.pitch, a.N, a.H, a.W, a.C, _vp(y), H, W, cx.stream)
        return y

    def import_grad(self, cx, i, d):
        if d is None:
            return None
        N, h, w, Cc, pitch = self.small
        scale = pop_pending_scale(d)       # a loss that left its mean-reduction factor to this kernel (functional.py)
        d = d.contiguous()
        g = cx.new(N, h, w, pitch)
        g.C = Cc
        L.call("s2r_upsample_bilinear_nchw_bwd_to_nhwc_scaled", _vp(d), N, Cc, d.shape[2], d.shape[3], g.vp(), pitch, h, w,
               _vp(scale), cx.stream)
        return g


class DecoderRun(RunBase):
    def __init__(self, mod, out_hw=None):
        self.out_hw = out_hw
        if out_hw is not None:      # Decoder.forward(..., size=...): fused up-sampling at the boundary
            self.export = lambda cx, i, a: UpsampledLogits.export(self, cx, i, a)
            self.import_grad = lambda cx, i, d: UpsampledLogits.import_grad(self, cx, i, d)
        lc = mod.last_conv
        self.low = ConvBNAct(mod.conv1, mod.bn1, L.ACT_RELU)
        self.c1 = ConvBNAct(lc[0], lc[1], L.ACT_RELU, drop_p=lc[3].p)
        self.c2 = ConvBNAct(lc[4], lc[5], L.ACT_RELU, drop_p=lc[7].p)
        self.cls = lc[8]

    def forward(self, cx, x, low):
        N, H, W = low.N, low.H, low.W
        cat = cx.new(N, H, W, 256 + 48)
        L.call("s2r_upsample_bilinear_nhwc", x.vp(), x.N, x.H, x.W, x.C, cat.vp(), H, W, cat.pitch, 0, cx.stream)
        self.low.forward(cx, low, out=cat.slice(256, 48))
        cx.tr('dec_cat', cat)
        y2 = cx.tr('dec_y2', self.c2.forward(cx, cx.tr('dec_y1', self.c1.forward(cx, cat))))
        ncls = self.cls.weight.shape[0]
        out = cx.new(N, H, W, round_up(ncls, 8), zero=True)
        out.C = ncls
        conv_fwd(cx, y2, self.cls.weight, out, bias=self.cls.bias)
        self.saved = (x, y2)
        return cx.tr('dec_logits', out)

    def backward(self, cx, douts, need=None):
        dlogits = douts[0] if isinstance(douts, tuple) else douts
        x, y2 = self.saved
        self.saved = None
        w, b = self.cls.weight, self.cls.bias
        if w.requires_grad:
            conv_wgrad(cx, y2, dlogits, w)
        if b is not None and b.requires_grad:
            bias_grad(cx, dlogits, b)
        dy2 = cx.new(y2.N, y2.H, y2.W, y2.C)
        conv_dgrad(cx, dlogits, w, dy2)
        dcat = self.c1.backward(cx, self.c2.backward(cx, dy2))
        dlow = self.low.backward(cx, dcat.slice(256, 48))
        dx = cx.new(x.N, x.H, x.W, x.C)
        L.call("s2r_upsample_bilinear_nhwc_bwd", dcat.vp(), dcat.pitch, 0, x.N, x.H, x.W, x.C, dcat.H, dcat.W,
               dx.vp(), cx.stream)
        return dx, dlow


class Decoder(nn.Module):
    def __init__(self, num_classes, backbone, BatchNorm):
        super().__init__()
        if backbone == 'resnet' or backbone == 'drn':
            low_level_inplanes = 256
        elif backbone == 'xception':
            low_level_inplanes = 128
        elif backbone == 'mobilenet':
            low_level_inplanes = 24
        else:
            raise NotImplementedError
        self.conv1 = nn.Conv2d(low_level_inplanes, 48, 1, bias=False)
        self.bn1 = BatchNorm(48)
        self.relu = nn.ReLU()
        self.last_conv = nn.Sequential(nn.Conv2d(304, 256, kernel_size=3, stride=1, padding=1, bias=False),
                                       BatchNorm(256),
                                       nn.ReLU(),
                                       nn.Dropout(0.5),
                                       nn.Conv2d(256, 256, kernel_size=3, stride=1, padding=1, bias=False),
                                       BatchNorm(256),
                                       nn.ReLU(),
                                       nn.Dropout(0.1),
                                       nn.Conv2d(256, num_classes, kernel_size=1, stride=1))
        init_reference_weights(self)
        self._s2r_has_sync_bn = bool(getattr(BatchNorm, "_s2r_sync", False))

    def forward(self, x, low_level_feat, size=None):
        """modeling/decoder.py:34-43.  size (extension, default None = the reference's behaviour): return
        F.interpolate(decoder(x, low), size=size, mode='bilinear', align_corners=True) -- what train.py:184,194,270
        wrap around the call -- with the up-sampling fused into the module's output conversion."""
        out_hw = None if size is None else (int(size[0]), int(size[1]))
        return call_module(self, lambda: DecoderRun(self, out_hw), (x, low_level_feat))


def build_decoder(num_classes, backbone, BatchNorm):
    return Decoder(num_classes, backbone, BatchNorm)
