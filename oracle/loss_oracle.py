"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement (torch fp64 on the CPU, autograd for the gradients) of the loss side of the mask logits
(SURVEY.md section 8f-4).  Paths relative to /root/reference/mask3d:

  matcher_costs        models/matcher.py:17-56, 99-160   (cost_class, cost_mask, cost_dice, C)
  hungarian            models/matcher.py:162              (scipy.optimize.linear_sum_assignment)
  mask_losses          models/criterion.py:27-72, 168-213 (sigmoid_ce_loss / dice_loss of the matched pairs)
  label_loss           models/criterion.py:141-166        (weighted cross entropy, unmatched queries -> class num_classes)
  set_criterion        models/criterion.py:237-281        (final output + every auxiliary decoder output)
  mask_scores          trainer/trainer.py:389-392

PINNED: tests/golden/loss_tiny.npz was written by the reference's UNMODIFIED HungarianMatcher + SetCriterion
(tests/golden/make_golden_loss.py imports them from /root/reference); tests/test_loss.py replays it through this file.
Only num_points == -1 (all points, the reference's configuration: conf/matcher/hungarian_matcher.yaml:5) and explicit row
subsets are restated; the random `randperm` subset itself is the caller's.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment

IGNORE_LABEL = 253  # matcher.py:107


def _bce_pos_neg(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    # matcher.py:52-53: BCE-with-logits against all-ones / all-zeros
    return F.softplus(-x), F.softplus(x)


def matcher_costs(pred_class: torch.Tensor, pred_mask: torch.Tensor, labels: torch.Tensor, masks: torch.Tensor,
                  rows: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """pred_class [Q, C+1] logits, pred_mask [N, Q] logits, labels [T], masks [T, N] bool.  matcher.py:99-160."""
    Q = pred_class.shape[0]
    T = labels.shape[0]
    prob = pred_class.double().softmax(-1)
    lab = labels.clone().long()
    ign = lab == IGNORE_LABEL
    lab[ign] = 0
    cost_class = -prob[:, lab] if T > 0 else torch.zeros((Q, 0), dtype=torch.float64)
    cost_class[:, ign] = -1.0
    x = pred_mask.double().T  # [Q, N]
    tgt = masks.double().reshape(T, x.shape[1])
    if rows is not None:
        x, tgt = x[:, rows], tgt[:, rows]
    hw = x.shape[1]
    pos, neg = _bce_pos_neg(x)
    cost_mask = (pos @ tgt.T + neg @ (1 - tgt).T) / hw  # matcher.py:55-56
    s = x.sigmoid()
    cost_dice = 1 - (2 * (s @ tgt.T) + 1) / (s.sum(-1)[:, None] + tgt.sum(-1)[None, :] + 1)  # matcher.py:28-32
    return dict(cost_class=cost_class, cost_mask=cost_mask, cost_dice=cost_dice)


def hungarian(C: torch.Tensor) -> Tuple[np.ndarray, np.ndarray]:
    i, j = linear_sum_assignment(C.detach().cpu().numpy())
    return i.astype(np.int64), j.astype(np.int64)


def mask_losses(pred_mask: torch.Tensor, masks: torch.Tensor, idx_q: Sequence[int], idx_t: Sequence[int],
                rows: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """criterion.py:176-206 for one batch item; num_masks is the item's own matched count (:190)."""
    idx_q = torch.as_tensor(idx_q, dtype=torch.long)
    idx_t = torch.as_tensor(idx_t, dtype=torch.long)
    x = pred_mask.double()[:, idx_q].T  # [M, N]
    tgt = masks[idx_t].double().reshape(len(idx_t), x.shape[1])
    if rows is not None:
        x, tgt = x[:, rows], tgt[:, rows]
    M = tgt.shape[0]
    if x.shape[1] == 0 or M == 0:  # criterion.py:197-203
        z = pred_mask.double().sum() * 0
        return z, z
    bce = F.binary_cross_entropy_with_logits(x, tgt, reduction="none")
    loss_mask = bce.mean(1).sum() / M
    s = x.sigmoid()
    loss_dice = (1 - (2 * (s * tgt).sum(-1) + 1) / (s.sum(-1) + tgt.sum(-1) + 1)).sum() / M
    return loss_mask, loss_dice


def label_loss(pred_class: torch.Tensor, labels: List[torch.Tensor], matches, num_classes: int, empty_weight: torch.Tensor):
    """criterion.py:141-166.  pred_class [B, Q, C+1]."""
    B, Q = pred_class.shape[:2]
    target = torch.full((B, Q), num_classes, dtype=torch.long)
    for b, (iq, it) in enumerate(matches):
        target[b, torch.as_tensor(iq, dtype=torch.long)] = labels[b][torch.as_tensor(it, dtype=torch.long)].long()
    return F.cross_entropy(pred_class.double().transpose(1, 2), target, weight=empty_weight.double())


def set_criterion(pred_class_all: List[torch.Tensor], pred_mask_all: List[List[torch.Tensor]], labels: List[torch.Tensor],
                  masks: List[torch.Tensor], num_classes: int, eos_coef: float, w_class: float, w_mask: float, w_dice: float):
    """criterion.py:237-281: the LAST entry of the lists is the final output (keys without suffix), the others are the
    auxiliary outputs `_mask_module_{i}`.  Returns (losses, matches per output)."""
    empty_weight = torch.ones(num_classes + 1, dtype=torch.float64)
    empty_weight[-1] = eos_coef
    losses, all_matches = {}, []
    order = [len(pred_class_all) - 1] + list(range(len(pred_class_all) - 1))
    for d in order:
        cls, pm = pred_class_all[d], pred_mask_all[d]
        matches = []
        for b in range(cls.shape[0]):
            c = matcher_costs(cls[b].detach(), pm[b].detach(), labels[b], masks[b])
            C = w_mask * c["cost_mask"] + w_class * c["cost_class"] + w_dice * c["cost_dice"]
            matches.append(hungarian(C.reshape(cls.shape[1], -1)))
        suffix = "" if d == len(pred_class_all) - 1 else f"_mask_module_{d}"
        losses["loss_ce" + suffix] = label_loss(cls, labels, matches, num_classes, empty_weight)
        lm = [mask_losses(pm[b], masks[b], *matches[b]) for b in range(cls.shape[0])]
        losses["loss_mask" + suffix] = torch.stack([a for a, _ in lm]).sum()
        losses["loss_dice" + suffix] = torch.stack([a for _, a in lm]).sum()
        all_matches.append((d, matches))
    return losses, dict(all_matches)


def mask_scores(pred_mask: torch.Tensor) -> torch.Tensor:
    """trainer.py:389-392 without the top-k column permutation: per query column."""
    x = pred_mask.double()
    hard = (x > 0).double()
    return (x.sigmoid() * hard).sum(0) / (hard.sum(0) + 1e-6)


def get_mask_and_scores(pred_class: torch.Tensor, pred_mask: torch.Tensor, num_queries: int, num_classes: int, topk_per_image: int = -1):
    """trainer/trainer.py:373-402, line by line (CPU)."""
    labels = torch.arange(num_classes).unsqueeze(0).repeat(num_queries, 1).flatten(0, 1)
    k = topk_per_image if topk_per_image != -1 else num_queries
    scores_per_query, topk_indices = pred_class.flatten(0, 1).topk(k, sorted=True)
    labels_per_query = labels[topk_indices]
    topk_indices = topk_indices // num_classes
    pm = pred_mask[:, topk_indices]
    result_pred_mask = (pm > 0).float()
    heatmap = pm.float().sigmoid()
    mask_scores_per_image = (heatmap * result_pred_mask).sum(0) / (result_pred_mask.sum(0) + 1e-6)
    return scores_per_query * mask_scores_per_image, result_pred_mask.bool(), labels_per_query, heatmap
