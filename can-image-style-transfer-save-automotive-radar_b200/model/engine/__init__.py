from .transfer_style import do_transfer_style
from .hr_transfer_style import do_hr_transfer_style

__all__ = ['do_transfer_style', 'do_hr_transfer_style']
