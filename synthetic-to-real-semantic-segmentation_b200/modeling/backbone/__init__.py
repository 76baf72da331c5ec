"""Backbone factory with the reference's signature (modeling/backbone/__init__.py:3-13).  The reference also names
resnet / xception / drn modules that are not in its tree (calling them is a NameError there); only MobileNetV2 exists."""
from .mobilenet import MobileNetV2

_BACKBONES = {'mobilenet': MobileNetV2}


def build_backbone(backbone, output_stride, BatchNorm):
    try:
        factory = _BACKBONES[backbone]
    except KeyError:
        raise NotImplementedError(backbone) from None
    return factory(output_stride, BatchNorm)
