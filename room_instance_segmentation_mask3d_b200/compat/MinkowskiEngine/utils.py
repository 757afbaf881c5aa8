from room_instance_segmentation_mask3d_b200.me_api import sparse_collate, sparse_quantize  # noqa: F401
