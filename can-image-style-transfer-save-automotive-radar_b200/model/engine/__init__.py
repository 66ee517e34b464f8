from .transfer_style import do_transfer_style, do_transfer_style_batch
from .hr_transfer_style import do_hr_transfer_style

__all__ = ['do_transfer_style', 'do_transfer_style_batch', 'do_hr_transfer_style']
