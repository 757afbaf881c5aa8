"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

Import-time shims that let the reference's UNMODIFIED model code (`/root/reference/mask3d/models`)
run on CPU on top of oracle/me_oracle.py:

  MinkowskiEngine (+ .MinkowskiOps, .MinkowskiPooling, .utils)  -> oracle.me_oracle
  torch_scatter.scatter_{mean,max,min}                           -> bound at mask3d.py:12,138-143, never called
  pointnet2._ext.furthest_point_sampling                         -> oracle.fps_oracle (pointnet2_utils.py:22-31,72)

Used only by tests/golden/make_golden.py (fixture generation in the build container, where
/root/reference exists) and by the optional reference-replay test.
"""
from __future__ import annotations

import sys
import types

import numpy as np
import torch

from . import fps_oracle, me_oracle


def install(reference_root: str = "/root/reference/mask3d") -> None:
    me = types.ModuleType("MinkowskiEngine")
    for name in dir(me_oracle):
        if not name.startswith("_"):
            setattr(me, name, getattr(me_oracle, name))
    ops = types.ModuleType("MinkowskiEngine.MinkowskiOps")
    ops.SparseTensor = me_oracle.SparseTensor
    ops.cat = me_oracle.cat
    pooling = types.ModuleType("MinkowskiEngine.MinkowskiPooling")
    for n in ("MinkowskiAvgPooling", "MinkowskiSumPooling", "MinkowskiAvgUnpooling", "MinkowskiPoolingTranspose", "MinkowskiGlobalPooling"):
        setattr(pooling, n, getattr(me_oracle, n))
    utils = types.ModuleType("MinkowskiEngine.utils")
    utils.sparse_collate = me_oracle.sparse_collate
    utils.sparse_quantize = me_oracle.sparse_quantize
    me.MinkowskiOps, me.MinkowskiPooling, me.utils = ops, pooling, utils
    me.__path__ = []  # mark as package
    sys.modules.update({"MinkowskiEngine": me, "MinkowskiEngine.MinkowskiOps": ops,
                        "MinkowskiEngine.MinkowskiPooling": pooling, "MinkowskiEngine.utils": utils})

    ts = types.ModuleType("torch_scatter")

    def _never(*a, **k):
        raise NotImplementedError("torch_scatter is bound but never called on the hot path (mask3d.py:138-143)")

    ts.scatter_mean = ts.scatter_max = ts.scatter_min = _never
    sys.modules["torch_scatter"] = ts

    p2 = types.ModuleType("pointnet2")
    ext = types.ModuleType("pointnet2._ext")

    def furthest_point_sampling(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        out = [fps_oracle.furthest_point_sample(xyz[b].detach().cpu().numpy(), npoint) for b in range(xyz.shape[0])]
        return torch.from_numpy(np.stack(out).astype(np.int32))

    ext.furthest_point_sampling = furthest_point_sampling
    p2._ext = ext
    p2.__path__ = []
    sys.modules["pointnet2"] = p2
    sys.modules["pointnet2._ext"] = ext

    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
