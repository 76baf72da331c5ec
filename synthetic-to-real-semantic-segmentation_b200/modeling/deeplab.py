"""DeepLabV3+ with the reference's constructor, attributes, parameter-group generators and
state_dict keys (modeling/deeplab.py:9-72).

forward(input[N,3,H,W] fp32) -> [N,num_classes,H,W] fp32.  Backbone, ASPP and decoder run
back-to-back on NHWC bf16 activations (no NCHW round trips between them); the final x4 bilinear
up-sampling (deeplab.py:31) writes the NCHW fp32 logits the callers expect.
"""
import torch.nn as nn

from ..runtime import RunBase, call_module
from .sync_batchnorm import SynchronizedBatchNorm2d
from .assp import build_aspp, ASPPRun
from .decoder import build_decoder, DecoderRun, UpsampledLogits
from .backbone import build_backbone
from .backbone.mobilenet import MobileNetV2Run

import torch


class DeepLabRun(UpsampledLogits, RunBase):
    """F.interpolate(x, size=input.size()[2:], mode='bilinear', align_corners=True) (deeplab.py:31) is fused with the
    NHWC bf16 -> NCHW fp32 boundary conversion (decoder.UpsampledLogits)."""
    raw_inputs = True   # see MobileNetV2Run

    def __init__(self, mod):
        self.backbone = MobileNetV2Run(mod.backbone)
        self.aspp = ASPPRun(mod.aspp)
        self.decoder = DecoderRun(mod.decoder)

    def forward(self, cx, x):
        self.out_hw = (x.H, x.W)
        high, low = self.backbone.forward(cx, x)
        return self.decoder.forward(cx, self.aspp.forward(cx, high), low)

    def backward(self, cx, douts, need=None):
        dx, dlow = self.decoder.backward(cx, douts)
        dhigh = self.aspp.backward(cx, dx)
        self.backbone.backward(cx, (dhigh, dlow))
        return None


class DeepLabConfusionRun(DeepLabRun):
    """Validation: the up-sampling at the module boundary, the argmax and the Evaluator's histogram are one launch on
    the decoder's low-resolution logits (utils.metrics.Evaluator.add_batch_lowres); the fp32 logits never exist."""

    def __init__(self, mod, target, evaluator):
        DeepLabRun.__init__(self, mod)
        self.target, self.evaluator = target, evaluator

    def export(self, cx, i, a):
        if tuple(self.target.shape[-2:]) != tuple(self.out_hw):
            raise ValueError("forward_confusion: label map %s, image %s" % (tuple(self.target.shape), self.out_hw))
        self.evaluator.add_batch_lowres(self.target, a.ptr, a.pitch, a.N, a.H, a.W, a.C, cx.stream)
        return torch.empty((a.N, 0), dtype=torch.float32, device=cx.device)


class DeepLab(nn.Module):
    def __init__(self, backbone='resnet', output_stride=16, num_classes=19, sync_bn=True, freeze_bn=False):
        super().__init__()
        if backbone == 'drn':
            output_stride = 8
        BatchNorm = SynchronizedBatchNorm2d if sync_bn == True else nn.BatchNorm2d  # noqa: E712
        self.backbone = build_backbone(backbone, output_stride, BatchNorm)
        self.aspp = build_aspp(backbone, output_stride, BatchNorm)
        self.decoder = build_decoder(num_classes, backbone, BatchNorm)
        self.freeze_bn = freeze_bn
        self._s2r_has_sync_bn = bool(sync_bn)

    def forward(self, input):
        return call_module(self, lambda: DeepLabRun(self), (input,))

    @torch.no_grad()
    def forward_confusion(self, input, target, evaluator):
        """val_adapt.py:126-135 in one call: `output = model(image)`, `pred = argmax(output, 1)`,
        `evaluator.add_batch(target, pred)` -- with the final x4 up-sampling (deeplab.py:31), the argmax and the
        histogram fused into one kernel on the decoder's low-resolution logits.  Counts are identical to
        evaluator.add_batch_logits(target, self(input)); eval mode only."""
        if self.training:
            raise RuntimeError("forward_confusion is the validation path: call model.eval() first")
        evaluator._ensure(input.device if input.device.type == "cuda" else None)
        call_module(self, lambda: DeepLabConfusionRun(self, target, evaluator), (input,))

    def _lr_params(self, modules):
        # deeplab.py:42-72: Conv2d (+ BatchNorm unless freeze_bn) parameters of the given sub-modules
        for root in modules:
            for _, m in root.named_modules():
                wanted = isinstance(m, nn.Conv2d) or \
                    (not self.freeze_bn and isinstance(m, nn.modules.batchnorm._BatchNorm))
                if wanted:
                    for p in m.parameters():
                        if p.requires_grad:
                            yield p

    def get_1x_lr_params(self):
        return self._lr_params([self.backbone])

    def get_10x_lr_params(self):
        return self._lr_params([self.aspp, self.decoder])
