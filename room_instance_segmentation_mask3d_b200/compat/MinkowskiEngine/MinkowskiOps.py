from room_instance_segmentation_mask3d_b200.me_api import SparseTensor, cat  # noqa: F401
