"""`pointnet2._ext` replacement: only furthest_point_sampling is reached (mask3d/models/mask3d.py:309-317 via
third_party/pointnet2/pointnet2_utils.py:72)."""
import torch

from room_instance_segmentation_mask3d_b200 import ops


def furthest_point_sampling(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    """xyz fp32 [B, N, 3] contiguous CUDA -> int32 [B, npoint]; same tie-breaking as the reference kernel."""
    assert xyz.is_cuda and xyz.dim() == 3 and xyz.shape[2] == 3
    B, N, _ = xyz.shape
    offsets = torch.arange(0, (B + 1) * N, N, dtype=torch.int32, device=xyz.device)
    return ops.furthest_point_sample(xyz.reshape(B * N, 3), offsets, B, npoint)
