"""Device-side input stage: the reference's `dataloders` transforms on uint8 batches resident in HBM."""
from .device_transforms import DeviceTrainTransform, DeviceValTransform, segmap_lut  # noqa: F401
