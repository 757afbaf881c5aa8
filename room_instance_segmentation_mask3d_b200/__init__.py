"""B200-native sparse-convolution / masked-attention hot path of Mask3D room instance segmentation.

`me_api`      MinkowskiEngine-compatible surface (drop-in boundary), `compat/` makes it importable as `MinkowskiEngine`
`ops`         functional wrappers over the C ABI (include/mk3d.h, libmk3d.so built from csrc/)
`engine`      coordinate manager: GPU hash maps, kernel maps, tensor-core tile plans
`mask3d_model` host-side mirror of Res16UNet34C + the Mask3D query decoder (state-dict compatible)
"""
import os
import sys

__version__ = "0.1.0"


def compat_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "compat")


def install_compat() -> None:
    """Make `import MinkowskiEngine`, `import torch_scatter`, `import pointnet2._ext` resolve to this package."""
    p = compat_path()
    if p not in sys.path:
        sys.path.insert(0, p)
