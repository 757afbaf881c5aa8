"""`import MinkowskiEngine` -> the B200 implementation (room_instance_segmentation_mask3d_b200.me_api).

Put `room_instance_segmentation_mask3d_b200/compat` on sys.path (or call
room_instance_segmentation_mask3d_b200.install_compat()) and the reference's mask3d code imports this
package in place of MinkowskiEngine 0.5.4.
"""
from room_instance_segmentation_mask3d_b200.me_api import *  # noqa: F401,F403
from room_instance_segmentation_mask3d_b200.me_api import __all__ as _all
from . import MinkowskiOps, MinkowskiPooling, utils  # noqa: F401

__version__ = "0.5.4+b200"
__all__ = list(_all)
