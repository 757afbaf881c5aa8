"""Generate tests/golden/loss_tiny.npz by running the reference's UNMODIFIED loss code.

Run in the build container only (it needs /root/reference):

    python tests/golden/make_golden_loss.py

What runs: `models.matcher.HungarianMatcher` and `models.criterion.SetCriterion` imported from /root/reference/mask3d
(pure PyTorch + scipy, importable once `detectron2` — imported at criterion.py:13-17, never called with num_points == -1 —
and the in-tree `datasets` package, shadowed here by the Hugging Face package of the same name, are put on sys.modules),
configured as the reference does (conf/matcher/hungarian_matcher.yaml: cost_class 2, cost_mask 5, cost_dice 2,
num_points -1; conf/loss/set_criterion.yaml: eos_coef 0.1, loss_names [labels, masks]; trainer/trainer.py:155-169,235-236
for the weights), on synthetic `ModelOutput` / `ItemTargets` objects of the reference's own classes.

The fixture pins: matcher cost matrices, the assignment, every loss of `SetCriterion.forward` (final + auxiliary decoder
outputs) and the gradient of the weighted loss sum w.r.t. the mask logits and class logits.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import shims  # noqa: E402

REF = "/root/reference/mask3d"


def install_loss_shims():
    shims.install(REF)
    mods = {}
    for name in ("detectron2", "detectron2.utils", "detectron2.projects", "detectron2.projects.point_rend"):
        m = types.ModuleType(name)
        m.__path__ = []
        mods[name] = m
    comm = types.ModuleType("detectron2.utils.comm")
    comm.get_world_size = lambda: 1
    pf = types.ModuleType("detectron2.projects.point_rend.point_features")
    pf.get_uncertain_point_coords_with_randomness = pf.point_sample = None  # bound at criterion.py:13-17, never called
    mods["detectron2.utils.comm"] = comm
    mods["detectron2.projects.point_rend.point_features"] = pf
    ds = types.ModuleType("datasets")  # the reference's own package, not the Hugging Face one
    ds.__path__ = [os.path.join(REF, "datasets")]
    mods["datasets"] = ds
    sys.modules.update(mods)


def synth_case(seed: int, sizes=(700, 413), n_targets=(5, 3), Q=25, D=3, num_classes=3):
    """Logits correlated with the targets (so that the assignment is meaningful), one overlapping pair of instances, one
    query far in the saturated range on both sides."""
    g = torch.Generator().manual_seed(seed)
    masks, labels, pred_mask_all, pred_class_all = [], [], [], []
    for n, T in zip(sizes, n_targets):
        inst = torch.div(torch.arange(n) * T, n, rounding_mode="floor")  # runs of rows, like lexicographic voxel order
        m = torch.stack([inst == t for t in range(T)]) if T > 0 else torch.zeros((0, n), dtype=torch.bool)
        if T >= 2:
            m[1, : n // 7] = True  # overlap with instance 0
        masks.append(m)
        labels.append(torch.randint(0, num_classes, (T,), generator=g))
    for d in range(D):
        pm = []
        for b, (n, T) in enumerate(zip(sizes, n_targets)):
            x = torch.randn(n, Q, generator=g) * 2.0
            perm = torch.randperm(Q, generator=g)
            for t in range(T):
                x[:, perm[t]] += (masks[b][t].float() * 2 - 1) * (1.5 + d)
            x[:, perm[-1]] = x[:, perm[-1]] * 20 - 30  # saturated: softplus / sigmoid tails
            x[: n // 3, perm[-2]] += 25
            pm.append(x)
        pred_mask_all.append(pm)
        pred_class_all.append(torch.randn(len(sizes), Q, num_classes + 1, generator=g))
    return masks, labels, pred_mask_all, pred_class_all


def main():
    install_loss_shims()
    from datasets.utils import ItemTargets  # noqa: E402  (reference)
    from models.criterion import SetCriterion  # noqa: E402  (reference, unmodified)
    from models.mask3d import ModelOutput  # noqa: E402  (reference)
    from models.matcher import HungarianMatcher, batch_dice_loss_jit, batch_sigmoid_ce_loss_jit  # noqa: E402

    num_classes, D = 3, 3
    masks, labels, pred_mask_all, pred_class_all = synth_case(seed=7, num_classes=num_classes, D=D)
    for pm in pred_mask_all:
        for x in pm:
            x.requires_grad_(True)
    for c in pred_class_all:
        c.requires_grad_(True)
    targets = [ItemTargets(labels[b], masks[b], torch.zeros(masks[b].shape[1], dtype=torch.long)) for b in range(len(masks))]
    matcher = HungarianMatcher(cost_class=2.0, cost_mask=5.0, cost_dice=2.0, num_points=-1)
    weight = {"loss_ce": matcher.cost_class, "loss_mask": matcher.cost_mask, "loss_dice": matcher.cost_dice}
    weight_dict = dict(weight)
    for d in range(D - 1):
        weight_dict.update({k + f"_mask_module_{d}": v for k, v in weight.items()})  # trainer.py:161-167
    crit = SetCriterion(num_classes=num_classes, matcher=matcher, weight_dict=weight_dict, eos_coef=0.1, loss_names=["labels", "masks"],
                        num_points=-1, oversample_ratio=3.0, importance_sample_ratio=0.75, class_weights=-1)
    out = ModelOutput(pred_class=pred_class_all[-1], pred_class_all_decoders=pred_class_all, pred_mask=pred_mask_all[-1],
                      pred_mask_all_decoders=pred_mask_all, sampled_coords=None, backbone_features=None)
    losses = crit(out, targets)
    total = sum(losses[k] * weight_dict[k] for k in losses)  # trainer.py:233-240
    total.backward()

    fx = {"num_classes": np.int64(num_classes), "D": np.int64(D), "B": np.int64(len(masks)), "total": total.detach().numpy()}
    for b in range(len(masks)):
        fx[f"masks_{b}"] = masks[b].numpy()
        fx[f"labels_{b}"] = labels[b].numpy()
    for d in range(D):
        fx[f"pred_class_{d}"] = pred_class_all[d].detach().numpy()
        fx[f"grad_class_{d}"] = pred_class_all[d].grad.numpy()
        matches = matcher(types.SimpleNamespace(pred_class=pred_class_all[d].detach(), pred_mask=[x.detach() for x in pred_mask_all[d]]), targets)
        for b in range(len(masks)):
            x = pred_mask_all[d][b]
            fx[f"pred_mask_{d}_{b}"] = x.detach().numpy()
            fx[f"grad_mask_{d}_{b}"] = x.grad.numpy()
            fx[f"match_q_{d}_{b}"] = matches[b][0].numpy()
            fx[f"match_t_{d}_{b}"] = matches[b][1].numpy()
            with torch.no_grad():
                fx[f"cost_mask_{d}_{b}"] = batch_sigmoid_ce_loss_jit(x.detach().T, masks[b].float()).numpy()
                fx[f"cost_dice_{d}_{b}"] = batch_dice_loss_jit(x.detach().T, masks[b].float()).numpy()
    fx["loss_names"] = np.array(sorted(losses))
    fx["loss_values"] = np.array([float(losses[k]) for k in sorted(losses)])

    # matcher alone on a case with an ignore label (253, matcher.py:107-117) and a scene without targets
    masks2, labels2, pm2, pc2 = synth_case(seed=11, sizes=(300, 150, 64), n_targets=(4, 0, 2), D=1, num_classes=num_classes)
    labels2[0][1] = 253
    targets2 = [ItemTargets(labels2[b], masks2[b], torch.zeros(masks2[b].shape[1], dtype=torch.long)) for b in range(3)]
    m2 = matcher(types.SimpleNamespace(pred_class=pc2[0], pred_mask=pm2[0]), targets2)
    fx["case2_pred_class"] = pc2[0].numpy()
    for b in range(3):
        fx[f"case2_masks_{b}"] = masks2[b].numpy()
        fx[f"case2_labels_{b}"] = labels2[b].numpy()
        fx[f"case2_pred_mask_{b}"] = pm2[0][b].numpy()
        fx[f"case2_match_q_{b}"] = m2[b][0].numpy()
        fx[f"case2_match_t_{b}"] = m2[b][1].numpy()

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "loss_tiny.npz")
    np.savez_compressed(path, **fx)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: round(float(v), 5) for k, v in losses.items()})


if __name__ == "__main__":
    main()
