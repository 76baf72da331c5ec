"""Domain classifier of the FCN-in-the-wild feature adaptation, with the reference's (misspelt)
module/class names and state_dict keys (modeling/domian.py:7-47)."""
import torch.nn as nn

from .. import _lib as L
from ..engine import ConvBNAct, conv_fwd, conv_dgrad, conv_wgrad, bias_grad, round_up
from ..runtime import RunBase, call_module, init_reference_weights


class DomainClassiferRun(RunBase):
    def __init__(self, mod):
        self.l1 = ConvBNAct(mod.DC_adnn1[0], mod.DC_adnn1[1], L.ACT_RELU, drop_p=mod.DC_adnn1[3].p)
        self.l2 = ConvBNAct(mod.DC_adnn2[0], mod.DC_adnn2[1], L.ACT_RELU, drop_p=mod.DC_adnn2[3].p)
        self.cls = mod.DC_adnn3

    def forward(self, cx, x):
        y2 = self.l2.forward(cx, self.l1.forward(cx, x))
        Cout = self.cls.weight.shape[0]
        out = cx.new(x.N, x.H, x.W, round_up(Cout, 8), zero=True)
        out.C = Cout
        conv_fwd(cx, y2, self.cls.weight, out, pad=1, bias=self.cls.bias)
        self.y2 = y2
        return out

    def backward(self, cx, douts, need=None):
        d = douts[0] if isinstance(douts, tuple) else douts
        y2, self.y2 = self.y2, None
        w, b = self.cls.weight, self.cls.bias
        if w.requires_grad:
            conv_wgrad(cx, y2, d, w, pad=1)
        if b is not None and b.requires_grad:
            bias_grad(cx, d, b)
        dy2 = cx.new(y2.N, y2.H, y2.W, y2.C)
        conv_dgrad(cx, d, w, dy2, pad=1)
        need_dx = True if need is None else bool(need[0])
        return self.l1.backward(cx, self.l2.backward(cx, dy2), need_dx=need_dx)


class DomainClassifer(nn.Module):
    def __init__(self, backbone, BatchNorm, level='high'):
        super().__init__()
        if backbone == 'mobilenet' and level == 'high':
            in_channel = 256
        else:
            raise NotImplementedError
        self.DC_adnn1 = nn.Sequential(nn.Conv2d(in_channel, 1024, kernel_size=1, stride=1, padding=0, bias=False),
                                      BatchNorm(1024),
                                      nn.ReLU(),
                                      nn.Dropout(0.5))
        self.DC_adnn2 = nn.Sequential(nn.Conv2d(1024, 1024, kernel_size=3, stride=1, padding=1, bias=False),
                                      BatchNorm(1024),
                                      nn.ReLU(),
                                      nn.Dropout(0.5))
        self.DC_adnn3 = nn.Conv2d(1024, 2, kernel_size=3, stride=1, padding=1, bias=True)
        init_reference_weights(self)
        self._s2r_has_sync_bn = bool(getattr(BatchNorm, "_s2r_sync", False))

    def forward(self, input):
        return call_module(self, lambda: DomainClassiferRun(self), (input,))


def build_domaincls(backbone, BatchNorm):
    return DomainClassifer(backbone, BatchNorm)
