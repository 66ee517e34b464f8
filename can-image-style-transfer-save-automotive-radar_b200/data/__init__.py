from .image_transform import ImageTransform
from .device_image_transform import DeviceImageTransform

__all__ = ['ImageTransform', 'DeviceImageTransform']
