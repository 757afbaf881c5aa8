from room_instance_segmentation_mask3d_b200.me_api import (  # noqa: F401
    MinkowskiAvgPooling, MinkowskiAvgUnpooling, MinkowskiGlobalPooling, MinkowskiPoolingTranspose, MinkowskiSumPooling)
