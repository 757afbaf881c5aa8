"""Generate tests/golden/*.npz by running the reference's UNMODIFIED model code.

Run in the build container only (it needs /root/reference):

    python tests/golden/make_golden.py

What runs: `models.Res16UNet34C` + `models.Mask3D` imported from /root/reference/mask3d
(reference: mask3d/models/res16unet.py, mask3d/models/mask3d.py), constructed as SURVEY.md section 5
describes (no Hydra), on top of the CPU oracle of MinkowskiEngine (oracle/me_oracle.py via
oracle/shims.py — ME 0.5.4 itself is not installable here).  Weights are the
name-keyed synthetic tensors of room_instance_segmentation_mask3d_b200/weights.py, so the
fixture stores only (name, shape) and the outputs.

The fixture pins: the reference's module composition (state-dict names/shapes, layer order,
decoder glue incl. the all-masked rule and padding), FPS tie-breaking, the fourier PE, and stock
nn.MultiheadAttention numerics.  It cannot pin ME's own kernels (absent) — see oracle header.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import shims  # noqa: E402

shims.install("/root/reference/mask3d")

import MinkowskiEngine as ME  # noqa: E402  (the oracle)
import utils.measure_runtime as measure_runtime  # noqa: E402  (reference)
from models.mask3d import Mask3D  # noqa: E402  (reference, unmodified)
from models.res16unet import Res16UNet34C  # noqa: E402  (reference, unmodified)

from room_instance_segmentation_mask3d_b200 import scenes, weights  # noqa: E402


def build_reference_model(num_queries=25):
    torch.manual_seed(1)
    bb = Res16UNet34C(in_channels=1, out_channels=1,
                      config=SimpleNamespace(bn_momentum=0.02, conv1_kernel_size=5, dilations=[1, 1, 1, 1]), out_fpn=True)
    m = Mask3D(config=SimpleNamespace(backbone=bb), hidden_dim=128, num_queries=num_queries, num_heads=8, dim_feedforward=1024,
               sample_sizes=[200, 800, 3200, 12800, 51200], shared_decoder=True, num_classes=1, num_decoders=3, dropout=0.0,
               pre_norm=False, positional_encoding_type="fourier", non_parametric_queries=True, normalize_pos_enc=True,
               use_level_embed=False, scatter_type="mean", hlevels=[0, 1, 2, 3], use_np_features=False, max_sample_size=False,
               random_queries=False, gauss_scale=1.0, random_query_both=False, random_normal=False)
    m.eval()
    weights.load_synth(m, seed=0)
    return m


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    m = build_reference_model()
    sd = m.state_dict()
    names = np.array(list(sd.keys()))
    shapes = np.array([",".join(str(d) for d in v.shape) for v in sd.values()])

    sc = [scenes.tiny_scene(0), scenes.tiny_scene(1, (20, 16, 10))]
    coords, feats = scenes.collate(sc)
    x = ME.SparseTensor(coordinates=torch.from_numpy(coords), features=torch.from_numpy(feats))
    measure_runtime.notify_start_item()
    with torch.no_grad():
        out = m(x, is_eval=True)
        bb_out, fmaps = m.backbone(ME.SparseTensor(features=torch.from_numpy(feats), coordinate_map_key=x.coordinate_map_key,
                                                   coordinate_manager=x.coordinate_manager))
    mgr = x.coordinate_manager
    level_coords = {}
    for lvl, ts in enumerate((1, 2, 4, 8, 16)):
        level_coords[f"coords_l{lvl}"] = mgr.get_coordinates(ME.CoordinateMapKey((ts,) * 3, "")).numpy()
    k1 = ME.CoordinateMapKey((1, 1, 1), "")
    k2 = ME.CoordinateMapKey((2, 2, 2), "")
    fix = dict(
        names=names, shapes=shapes, coords=coords, feats=feats,
        backbone_out=bb_out.F.numpy(),
        fmap_sums=np.array([[float(f.F.double().sum()), float(f.F.double().abs().sum())] for f in fmaps]),
        fmap0=fmaps[0].F.numpy(), fmap3=fmaps[3].F.numpy(),
        pred_class=np.stack([p.numpy() for p in out.pred_class_all_decoders]),
        pred_mask_first=torch.cat(out.pred_mask_all_decoders[0]).numpy(),
        pred_mask_mid=torch.cat(out.pred_mask_all_decoders[6]).numpy(),
        pred_mask_last=torch.cat(out.pred_mask).numpy(),
        sampled_coords=out.sampled_coords,
        nbr3_l0=mgr.neighbor_table(k1, k1, 3, 1), nbr2_l0_l1=mgr.neighbor_table(k1, k2, 2, 1),
        nbr5_l0_rowsum=mgr.neighbor_table(k1, k1, 5, 1).astype(np.int64).sum(1),
        **level_coords,
    )
    path = os.path.join(out_dir, "mask3d_tiny2.npz")
    np.savez_compressed(path, **fix)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
