from . import mobilenet


def build_backbone(backbone, output_stride, BatchNorm):
    # modeling/backbone/__init__.py:3-13 of the reference names resnet/xception/drn modules that
    # are not in its tree; only MobileNetV2 exists.
    if backbone == 'mobilenet':
        return mobilenet.MobileNetV2(output_stride, BatchNorm)
    raise NotImplementedError
