"""Boundary proof, run as a SUBPROCESS by tests/test_gpu_boundary.py / tests/test_cpu_contract.py (a clean interpreter: the
reference's top-level package names `models`, `utils`, `third_party` must not collide with anything pytest has imported).

Puts room_instance_segmentation_mask3d_b200/compat (the drop-in `MinkowskiEngine`, `pointnet2._ext`, `torch_scatter`) and the
reference's OWN, UNMODIFIED mask3d/ tree on sys.path, builds `models.Res16UNet34C` + `models.Mask3D` exactly as
tests/golden/make_golden.py does (reference: mask3d/models/mask3d.py:73-259, res16unet.py:327-391), and

  --cpu   checks import + construction + state-dict names/shapes against the golden fixture (no CUDA needed);
  --gpu   loads the synthetic weights, runs the reference's `Mask3D.forward(x, is_eval=True)` (mask3d.py:280-518) on cuda:0 in
          fp32 parity mode and compares with tests/golden/mask3d_tiny2.npz (written by the same reference code on the CPU oracle);
          then a backward pass through the reference graph (autograd through the me_api Functions) compared with
          tests/golden/mask3d_tiny2_grads.npz.

Prints one JSON line."""
import json
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mode, ref_root = sys.argv[1], sys.argv[2]
sys.path[:0] = [os.path.join(ROOT, "room_instance_segmentation_mask3d_b200", "compat"), ROOT, ref_root]

import numpy as np  # noqa: E402
import torch  # noqa: E402

import MinkowskiEngine as ME  # noqa: E402  (compat -> room_instance_segmentation_mask3d_b200.me_api)

assert "room_instance_segmentation_mask3d_b200" in ME.SparseTensor.__module__, "MinkowskiEngine did not resolve to the drop-in"
import utils.measure_runtime as measure_runtime  # noqa: E402  (reference)
from models.mask3d import Mask3D  # noqa: E402  (reference, unmodified)
from models.res16unet import Res16UNet34C  # noqa: E402  (reference, unmodified)

assert os.path.abspath(sys.modules["models.mask3d"].__file__).startswith(os.path.abspath(ref_root)), "models.mask3d is not the reference's file"

from room_instance_segmentation_mask3d_b200 import ops, weights  # noqa: E402


def build_reference_model(num_queries=25):
    torch.manual_seed(1)
    bb = Res16UNet34C(in_channels=1, out_channels=1,
                      config=SimpleNamespace(bn_momentum=0.02, conv1_kernel_size=5, dilations=[1, 1, 1, 1]), out_fpn=True)
    return Mask3D(config=SimpleNamespace(backbone=bb), hidden_dim=128, num_queries=num_queries, num_heads=8, dim_feedforward=1024,
                  sample_sizes=[200, 800, 3200, 12800, 51200], shared_decoder=True, num_classes=1, num_decoders=3, dropout=0.0,
                  pre_norm=False, positional_encoding_type="fourier", non_parametric_queries=True, normalize_pos_enc=True,
                  use_level_embed=False, scatter_type="mean", hlevels=[0, 1, 2, 3], use_np_features=False, max_sample_size=False,
                  random_queries=False, gauss_scale=1.0, random_query_both=False, random_normal=False)


def grad_loss(out):
    """scalar the gradient fixture differentiates (same expression in tests/golden/make_golden_grads.py)"""
    loss = 0.0
    for pm, pc in zip(out.pred_mask_all_decoders, out.pred_class_all_decoders):
        loss = loss + torch.cat(pm).square().mean() + pc.square().mean()
    return loss


def grad_signature(model):
    """per parameter: (l2 norm, sum, dot with a fixed cosine pattern) of its gradient — 3 numbers instead of 39.6 M"""
    sig = {}
    for n, p in model.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.detach().double().cpu().reshape(-1)
        r = torch.cos(torch.arange(g.numel(), dtype=torch.float64) * 0.37 + 0.11)
        sig[n] = [float(g.norm()), float(g.sum()), float((g * r).sum())]
    return sig


def main():
    fix = np.load(os.path.join(ROOT, "tests", "golden", "mask3d_tiny2.npz"))
    m = build_reference_model()
    sd = m.state_dict()
    res = {"mode": mode, "n_params": len(sd)}
    want = dict(zip([str(n) for n in fix["names"]], [str(s) for s in fix["shapes"]]))
    got = {k: ",".join(str(d) for d in v.shape) for k, v in sd.items()}
    res["state_dict_matches_golden"] = got == want
    if mode == "--gpu":
        assert torch.cuda.is_available()
        m.eval()
        weights.load_synth(m, seed=0)
        m = m.cuda()
        ops.set_compute_mode("fp32")
        coords, feats = torch.from_numpy(fix["coords"]).cuda(), torch.from_numpy(fix["feats"]).cuda()
        x = ME.SparseTensor(coordinates=coords, features=feats)
        measure_runtime.notify_start_item()
        with torch.no_grad():
            out = m(x, is_eval=True)
            bb_out, fmaps = m.backbone(ME.SparseTensor(features=feats, coordinate_map_key=x.coordinate_map_key,
                                                       coordinate_manager=x.coordinate_manager))
        rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))
        res["backbone_rel_err"] = rel(bb_out.F.float().cpu().numpy(), fix["backbone_out"])
        res["fmap0_rel_err"] = rel(fmaps[0].F.float().cpu().numpy(), fix["fmap0"])
        res["pred_class_rel_err"] = rel(np.stack([p.float().cpu().numpy() for p in out.pred_class_all_decoders]), fix["pred_class"])
        res["pred_mask_first_rel_err"] = rel(torch.cat(out.pred_mask_all_decoders[0]).float().cpu().numpy(), fix["pred_mask_first"])
        res["pred_mask_mid_rel_err"] = rel(torch.cat(out.pred_mask_all_decoders[6]).float().cpu().numpy(), fix["pred_mask_mid"])
        res["pred_mask_last_rel_err"] = rel(torch.cat(out.pred_mask).float().cpu().numpy(), fix["pred_mask_last"])
        sc = out.sampled_coords
        sc = sc.cpu().numpy() if torch.is_tensor(sc) else np.asarray(sc)
        res["sampled_coords_equal"] = bool(np.array_equal(sc, fix["sampled_coords"]))
        mgr = x.coordinate_manager
        res["level_coords_equal"] = all(
            np.array_equal(mgr.get_coordinates(ME.CoordinateMapKey((ts,) * 3, "")).cpu().numpy(), fix[f"coords_l{lvl}"])
            for lvl, ts in enumerate((1, 2, 4, 8, 16)))
        # backward through the reference graph (autograd through the drop-in's Functions)
        gpath = os.path.join(ROOT, "tests", "golden", "mask3d_tiny2_grads.npz")
        for p in m.parameters():
            p.grad = None
        measure_runtime.notify_start_item()
        out = m(ME.SparseTensor(coordinates=coords, features=feats), is_eval=True)
        grad_loss(out).backward()
        torch.cuda.synchronize()
        sig = grad_signature(m)
        res["params_with_grad"] = len(sig)
        res["params_without_grad"] = sorted(n for n, p in m.named_parameters() if p.grad is None)
        res["grads_finite"] = all(np.isfinite(v).all() for v in sig.values())
        if os.path.exists(gpath):
            g = np.load(gpath)
            names = [str(n) for n in g["names"]]
            worst, worst_name = 0.0, ""
            for n, w in zip(names, g["sig"]):
                if n not in sig:
                    worst, worst_name = float("inf"), n + " (missing)"
                    break
                e = max(abs(sig[n][0] - w[0]), abs(sig[n][2] - w[2])) / max(w[0], 1e-12)  # norm + projection error relative to the norm
                if e > worst:
                    worst, worst_name = float(e), n
            res["grad_worst_rel_err"], res["grad_worst_param"], res["grad_params_compared"] = worst, worst_name, len(names)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
