from .vgg import VGG
from .gram_matrix import GramMatrix
from .gram_mse_loss import GramMSELoss
from .style_transfer import StyleTransfer

__all__ = ['VGG', 'GramMatrix', 'GramMSELoss', 'StyleTransfer']
