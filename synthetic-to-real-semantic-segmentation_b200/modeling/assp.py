"""ASPP head with the reference's names and state_dict keys (modeling/assp.py:7-96).

Four parallel conv(+BN+ReLU) branches (1x1 and three dilated 3x3) and the image-pooling branch
write straight into channel slices of one NHWC concat buffer (no torch.cat copy), followed by the
1x1 1280->256 projection, BN, ReLU and Dropout(0.5).
"""
import torch.nn as nn

from .. import _lib as L
from ..engine import ConvBNAct
from ..runtime import RunBase, call_module, init_reference_weights


class _ASPPModule(nn.Module):
    def __init__(self, inplanes, planes, kernel_size, padding, dilation, BatchNorm):
        super().__init__()
        self.atrous_conv = nn.Conv2d(inplanes, planes, kernel_size=kernel_size, stride=1, padding=padding,
                                     dilation=dilation, bias=False)
        self.bn = BatchNorm(planes)
        self.relu = nn.ReLU()
        init_reference_weights(self)

    def forward(self, x):
        raise L.S2RError("_ASPPModule is executed by its ASPP parent on the fused kernels")


class ASPPRun(RunBase):
    def __init__(self, mod):
        self.branches = [ConvBNAct(b.atrous_conv, b.bn, L.ACT_RELU)
                         for b in (mod.aspp1, mod.aspp2, mod.aspp3, mod.aspp4)]
        self.pool = ConvBNAct(mod.global_avg_pool[1], mod.global_avg_pool[2], L.ACT_RELU)
        self.proj = ConvBNAct(mod.conv1, mod.bn1, L.ACT_RELU, drop_p=mod.dropout.p)

    def forward(self, cx, x):
        N, H, W = x.N, x.H, x.W
        cat = cx.new(N, H, W, 1280)
        for i, b in enumerate(self.branches):
            b.forward(cx, x, out=cat.slice(256 * i, 256))
        pooled = cx.new(N, 1, 1, x.C)
        L.call("s2r_avgpool_nhwc", x.vp(), N, H * W, x.C, x.pitch, 0, 1.0 / (H * W), pooled.vp(), None, cx.stream)
        y5 = self.pool.forward(cx, pooled)
        # bilinear resize of a 1x1 map with align_corners=True is a broadcast (assp.py:71)
        L.call("s2r_broadcast_nhwc", y5.vp(), N, H * W, 256, 1.0, 0, cat.slice(1024, 256).vp(), cat.pitch, 0,
               cx.stream)
        self.shape = (N, H, W, x.C)
        cx.tr('aspp_cat', cat)
        return cx.tr('aspp_out', self.proj.forward(cx, cat))

    def backward(self, cx, douts, need=None):
        dout = douts[0] if isinstance(douts, tuple) else douts
        N, H, W, Cin = self.shape
        dcat = self.proj.backward(cx, dout)
        dx = None
        for i, b in enumerate(self.branches):
            dx = b.backward(cx, dcat.slice(256 * i, 256), dx=dx, dx_accumulate=dx is not None)
        dy5 = cx.new(N, 1, 1, 256)
        L.call("s2r_avgpool_nhwc", dcat.slice(1024, 256).vp(), N, H * W, 256, dcat.pitch, 0, 1.0, dy5.vp(), None,
               cx.stream)
        dpooled = self.pool.backward(cx, dy5)
        L.call("s2r_broadcast_nhwc", dpooled.vp(), N, H * W, Cin, 1.0 / (H * W), 1, dx.vp(), dx.pitch, 0, cx.stream)
        return dx


class ASPP(nn.Module):
    def __init__(self, backbone, output_stride, BatchNorm):
        super().__init__()
        inplanes = {'drn': 512, 'mobilenet': 320}.get(backbone, 2048)                 # assp.py:37-42
        dilations = {16: (1, 6, 12, 18), 8: (1, 12, 24, 36)}.get(output_stride)       # assp.py:43-48
        if dilations is None:
            raise NotImplementedError
        # aspp1 (1x1) and aspp2..4 (3x3, padding = dilation), registered -- and initialised -- in this order
        for k, d in enumerate(dilations, start=1):
            ksize = 1 if k == 1 else 3
            setattr(self, 'aspp%d' % k, _ASPPModule(inplanes, 256, ksize, padding=0 if k == 1 else d, dilation=d,
                                                    BatchNorm=BatchNorm))
        self.global_avg_pool = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)),
                                             nn.Conv2d(inplanes, 256, 1, stride=1, bias=False),
                                             BatchNorm(256),
                                             nn.ReLU())
        self.conv1 = nn.Conv2d(1280, 256, 1, bias=False)
        self.bn1 = BatchNorm(256)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(0.5)
        init_reference_weights(self)
        self._s2r_has_sync_bn = bool(getattr(BatchNorm, "_s2r_sync", False))

    def forward(self, x):
        return call_module(self, lambda: ASPPRun(self), (x,))


def build_aspp(backbone, output_stride, BatchNorm):
    return ASPP(backbone, output_stride, BatchNorm)
