"""B200-native hot path of haofengsiji/synthetic-to-real-semantic-segmentation.

Same module API as the reference (modeling.deeplab.DeepLab, modeling.discriminator.FCDiscriminator,
modeling.domian.DomainClassifer, utils.loss.SegmentationLosses / DomainLosses,
utils.metrics.Evaluator), every forward/backward executed by the sm_100a kernels in csrc/ through
the C ABI of include/s2r_b200.h.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
