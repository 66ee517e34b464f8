from .image_transform import ImageTransform

__all__ = ['ImageTransform']
