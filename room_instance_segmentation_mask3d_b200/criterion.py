"""Loss side of the mask logits: drop-in mirrors of the reference's `HungarianMatcher`
(mask3d/models/matcher.py:64-201) and `SetCriterion` (mask3d/models/criterion.py:93-290) on top of libmk3d.so.

Same constructor arguments, same `forward(model_output, targets)` contract, same loss-dict keys.  What changes is the work:

* targets are packed ONCE per step into per-voxel bit words (`pack_targets`, `mk_target_bits`);
* every decoder output is read once by `mk_mask_costs`, which yields the matcher's `cost_mask` / `cost_dice` matrices for all
  scenes of the batch (and the post-processing mask score of trainer/trainer.py:389-392);
* with `num_points == -1` (the reference's configuration) the cost matrices of all 13 decoder outputs reach the host in ONE
  copy before the 13 x B Hungarian assignments (scipy, as in the reference) run — the reference synchronises per output and
  per scene;
* the matched mask losses are gathers from those matrices (criterion.py:195-196 recomputes exactly cost_mask[q_i, t_i] and
  cost_dice[q_i, t_i]); their backward is one elementwise kernel (`mk_mask_loss_bwd`).

No CPU fallback: without a CUDA device every entry point raises.  The classification loss is the reference's own
`F.cross_entropy` on the [B, C+1, Q] logits (a few hundred numbers).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment
from torch import nn

from . import lib

IGNORE_LABEL = 253  # matcher.py:107


def _get(o, name):
    return o[name] if isinstance(o, dict) else getattr(o, name)


class PackedTargets:
    """Per-voxel target bit words of one batch + the host-side copies the matcher needs."""

    def __init__(self, bits, scene_offsets, sizes, n_targets, labels_host, labels_dev, ignore_dev):
        self.bits, self.scene_offsets, self.sizes, self.n_targets = bits, scene_offsets, sizes, n_targets
        self.labels_host, self.labels_dev, self.ignore_dev = labels_host, labels_dev, ignore_dev
        self.W = bits.shape[1]
        self.B = len(sizes)
        self.max_rows = max(sizes) if sizes else 0


def pack_targets(targets: Sequence, sizes: Sequence[int], device) -> PackedTargets:
    """targets: list of ItemTargets-like objects (`instances_labels [T]`, `instances_masks [T, N]`, datasets/utils.py:149-160);
    sizes: voxels per scene (the row counts of the mask logits)."""
    lib.require_cuda()
    device = torch.device(device)
    B = len(targets)
    n_targets = [int(_get(t, "instances_labels").shape[0]) for t in targets]
    W = max(1, (max(n_targets, default=0) + 31) // 32)
    total = int(sum(sizes))
    bits = torch.empty((total, W), dtype=torch.int32, device=device)
    offs = np.zeros(B + 1, dtype=np.int64)
    offs[1:] = np.cumsum(sizes)
    assert offs[-1] < 2 ** 31, "scene offsets are int32"
    st = lib.stream()
    keep = []
    for b, t in enumerate(targets):
        m = _get(t, "instances_masks")
        if n_targets[b] == 0:
            bits[offs[b]:offs[b + 1]].zero_()
            continue
        assert m.dim() == 2 and m.shape[1] == sizes[b], f"instances_masks of item {b}: expected [T, {sizes[b]}], got {tuple(m.shape)}"
        m = m.to(device=device)
        m8 = (m if m.dtype == torch.bool else (m != 0)).contiguous().view(torch.uint8)
        keep.append(m8)
        lib.call("mk_target_bits", lib.p(m8), n_targets[b], int(sizes[b]), W, bits.data_ptr() + int(offs[b]) * W * 4, st)
    Tp = 32 * W
    labels_host = [_get(t, "instances_labels").detach().cpu().long() for t in targets]
    lab = torch.zeros((B, Tp), dtype=torch.long)
    ign = torch.zeros((B, Tp), dtype=torch.bool)
    for b, l in enumerate(labels_host):
        ign[b, : len(l)] = l == IGNORE_LABEL
        lab[b, : len(l)] = torch.where(l == IGNORE_LABEL, torch.zeros_like(l), l)
    so = torch.from_numpy(offs.astype(np.int32)).to(device)
    return PackedTargets(bits, so, [int(s) for s in sizes], n_targets, labels_host, lab.to(device), ign.to(device))


def _batched_logits(pm, sizes) -> torch.Tensor:
    """A decoder output's mask logits as ONE [sum N, Q] fp32 tensor: passed as such (fused model), or a list of per-scene
    tensors (the reference's `pred_mask`), which is used without a copy when the items are the consecutive row blocks of one
    tensor (e.g. `torch.split` views)."""
    if isinstance(pm, torch.Tensor):
        return pm
    base = pm[0]._base if len(pm) and pm[0]._base is not None else None
    if base is not None and base.dim() == 2 and base.is_contiguous() and base.shape[0] == sum(sizes):
        off, ok = 0, True
        for t, n in zip(pm, sizes):
            ok = ok and t._base is base and t.shape[0] == n and t.shape[1] == base.shape[1] and t.is_contiguous() \
                and t.storage_offset() == base.storage_offset() + off * base.shape[1]
            off += n
        if ok:
            return base
    return torch.cat(list(pm), dim=0)


def _sizes_of(pm, targets) -> List[int]:
    if isinstance(pm, torch.Tensor):
        sizes = []
        for t in targets:
            m = _get(t, "instances_masks")
            assert m.dim() == 2, "batched logits need [T, N] instance masks (or pass per-scene logits) to find the scene sizes"
            sizes.append(int(m.shape[1]))
        return sizes
    return [int(t.shape[0]) for t in pm]


class MaskStats:
    __slots__ = ("cost_mask", "cost_dice", "dice_den", "mask_score", "n_rows", "row_select")


def mask_costs(logits: torch.Tensor, tg: PackedTargets, row_select: Optional[torch.Tensor] = None) -> MaskStats:
    """One pass over fp32 logits [sum N, Q] -> cost_mask / cost_dice / dice_den [B, Q, 32 W], mask_score [B, Q], n_rows [B]."""
    lib.require_cuda()
    x = logits.detach()
    if x.dtype != torch.float32:
        x = x.float()
    x = x.contiguous()
    n, Q = x.shape
    assert n == tg.bits.shape[0], f"logits have {n} rows, targets were packed for {tg.bits.shape[0]}"
    B, W = tg.B, tg.W
    dev = x.device
    out = torch.empty((3, B, Q, 32 * W), dtype=torch.float32, device=dev)
    score = torch.empty((B, Q), dtype=torch.float32, device=dev)
    n_rows = torch.empty((B,), dtype=torch.float32, device=dev)
    wsb = lib.query("mk_mask_costs_workspace_bytes", B, tg.max_rows, Q, W)
    ws = lib.scratch(wsb, dev)
    lib.call("mk_mask_costs", lib.p(x), n, Q, lib.p(tg.bits), W, lib.p(row_select), lib.p(tg.scene_offsets), B, tg.max_rows,
             out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), lib.p(score), lib.p(n_rows), lib.p(ws), ws.numel(), lib.stream())
    s = MaskStats()
    s.cost_mask, s.cost_dice, s.dice_den, s.mask_score, s.n_rows, s.row_select = out[0], out[1], out[2], score, n_rows, row_select
    return s


def mask_scores(pred_mask: torch.Tensor) -> torch.Tensor:
    """Per-query mask score of the post-processing (trainer/trainer.py:389-392, before the top-k column permutation):
    sum_n sigmoid(x) [x > 0] / (sum_n [x > 0] + 1e-6) for logits [N, Q] of ONE scene."""
    lib.require_cuda()
    n = int(pred_mask.shape[0])
    dev = pred_mask.device
    tg = PackedTargets(torch.zeros((n, 1), dtype=torch.int32, device=dev), torch.tensor([0, n], dtype=torch.int32, device=dev), [n], [0],
                       [], None, None)
    return mask_costs(pred_mask, tg).mask_score[0]


def _row_subset(tg: PackedTargets, num_points: float, device) -> torch.Tensor:
    """The reference's per-item random subset (matcher.py:125-127, criterion.py:181-182): torch.randperm(N)[: int(num_points N)],
    drawn per scene in scene order from the global generator of the device, as a uint8 row mask."""
    sel = torch.zeros(sum(tg.sizes), dtype=torch.uint8, device=device)
    off = 0
    for n in tg.sizes:
        idx = torch.randperm(n, device=device)[: int(num_points * n)]
        sel[idx + off] = 1
        off += n
    return sel


class HungarianMatcher(nn.Module):
    """Mirror of models.matcher.HungarianMatcher (matcher.py:64-201)."""

    def __init__(self, cost_class: float = 1, cost_mask: float = 1, cost_dice: float = 1, num_points: int = 0):
        super().__init__()
        self.cost_class, self.cost_mask, self.cost_dice = cost_class, cost_mask, cost_dice
        assert cost_class != 0 or cost_mask != 0 or cost_dice != 0, "all costs cant be 0"
        self.num_points = num_points

    def cost_matrix(self, pred_class: torch.Tensor, stats: MaskStats, tg: PackedTargets) -> torch.Tensor:
        """C [B, Q, 32 W] on the device (matcher.py:103-160); columns >= T of a scene are padding."""
        prob = pred_class.detach().float().softmax(-1)  # [B, Q, C+1]
        Q = prob.shape[1]
        cc = -prob.gather(2, tg.labels_dev[:, None, :].expand(-1, Q, -1))
        cc = torch.where(tg.ignore_dev[:, None, :], torch.full_like(cc, -1.0), cc)  # matcher.py:117
        return self.cost_mask * stats.cost_mask + self.cost_class * cc + self.cost_dice * stats.cost_dice

    @staticmethod
    def assign(C_host: torch.Tensor, tg: PackedTargets):
        """Hungarian assignment per scene on the host copy of C [B, Q, Tp] (matcher.py:160-162)."""
        assert not torch.isnan(C_host).any(), "out mask must not be NaN"  # matcher.py:120
        out = []
        for b in range(tg.B):
            i, j = linear_sum_assignment(C_host[b, :, : tg.n_targets[b]].numpy())
            out.append((torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)))
        return out

    @torch.no_grad()
    def memory_efficient_forward(self, model_output, targets, packed: Optional[PackedTargets] = None, return_stats: bool = False):
        pm = _get(model_output, "pred_mask")
        pred_class = _get(model_output, "pred_class")
        if packed is None:
            packed = pack_targets(targets, _sizes_of(pm, targets), pred_class.device)
        logits = _batched_logits(pm, packed.sizes)
        sel = None if self.num_points == -1 else _row_subset(packed, self.num_points, logits.device)
        stats = mask_costs(logits, packed, sel)
        C = self.cost_matrix(pred_class, stats, packed).cpu()
        matches = self.assign(C, packed)
        return (matches, stats) if return_stats else matches

    @torch.no_grad()
    def forward(self, model_output, targets):
        return self.memory_efficient_forward(model_output, targets)

    def __repr__(self, _repr_indent=4):
        body = [f"cost_class: {self.cost_class}", f"cost_mask: {self.cost_mask}", f"cost_dice: {self.cost_dice}"]
        return "\n".join(["Matcher " + self.__class__.__name__] + [" " * _repr_indent + line for line in body])


class _MaskLossFn(torch.autograd.Function):
    """(loss_mask, loss_dice) of the matched pairs of one decoder output, summed over the scenes of the batch
    (criterion.py:168-213).  Forward = gathers from the cost matrices; backward = mk_mask_loss_bwd."""

    @staticmethod
    def forward(ctx, logits, tg: PackedTargets, stats: MaskStats, match: torch.Tensor, inv_m: torch.Tensor):
        valid = match >= 0
        m = match.long().clamp(min=0).unsqueeze(-1)
        live = (valid & (stats.n_rows > 0)[:, None]).float()  # no points, no loss (criterion.py:197-203)
        cm = stats.cost_mask.gather(2, m).squeeze(-1) * live
        cd = stats.cost_dice.gather(2, m).squeeze(-1) * live
        den = stats.dice_den.gather(2, m).squeeze(-1)
        ctx.tg, ctx.row_select = tg, stats.row_select
        ctx.save_for_backward(logits, match, inv_m, cd, den, stats.n_rows, live)
        return torch.stack([(cm.sum(1) * inv_m).sum(), (cd.sum(1) * inv_m).sum()])

    @staticmethod
    def backward(ctx, g):
        logits, match, inv_m, cd, den, n_rows, live = ctx.saved_tensors
        tg = ctx.tg
        g = g.float()
        a = g[0] * inv_m[:, None] / n_rows.clamp(min=1.0)[:, None] * live
        c1 = g[1] * inv_m[:, None] * (1.0 - cd) / den * live  # num / den^2 = (1 - cost_dice) / den
        c2 = 2.0 * g[1] * inv_m[:, None] / den * live
        coef = torch.stack([a, c1, c2], dim=-1).contiguous()
        x = logits.detach()
        x = (x if x.dtype == torch.float32 else x.float()).contiguous()
        grad = torch.empty_like(x)
        n, Q = x.shape
        lib.call("mk_mask_loss_bwd", lib.p(x), n, Q, lib.p(tg.bits), tg.W, lib.p(ctx.row_select), lib.p(tg.scene_offsets), tg.B,
                 tg.max_rows, lib.p(match), lib.p(coef), lib.p(grad), lib.stream())
        return grad.to(logits.dtype), None, None, None, None


class SetCriterion(nn.Module):
    """Mirror of models.criterion.SetCriterion (criterion.py:93-290): same arguments, same keys in the returned dict
    (`loss_ce`, `loss_mask`, `loss_dice` + `_mask_module_{i}` for the auxiliary decoder outputs)."""

    def __init__(self, num_classes, matcher, weight_dict, eos_coef, loss_names, num_points, oversample_ratio, importance_sample_ratio,
                 class_weights):
        super().__init__()
        self.num_classes, self.class_weights, self.matcher, self.weight_dict = num_classes, class_weights, matcher, weight_dict
        self.eos_coef, self.loss_names = eos_coef, loss_names
        empty_weight = torch.ones(self.num_classes + 1)
        empty_weight[-1] = self.eos_coef
        if self.class_weights != -1:
            assert len(self.class_weights) == self.num_classes, "CLASS WEIGHTS DO NOT MATCH"
            empty_weight[:-1] = torch.tensor(self.class_weights)
        self.register_buffer("empty_weight", empty_weight)
        self.num_points, self.oversample_ratio, self.importance_sample_ratio = num_points, oversample_ratio, importance_sample_ratio

    # -- per-output pieces -------------------------------------------------------------------------------------------
    def _target_classes(self, matches, tg: PackedTargets, Q: int) -> torch.Tensor:
        tc = torch.full((tg.B, Q), self.num_classes, dtype=torch.int64)  # criterion.py:152-158
        for b, (iq, it) in enumerate(matches):
            tc[b, iq] = tg.labels_host[b][it]
        return tc

    @staticmethod
    def _match_table(matches, tg: PackedTargets, Q: int):
        m = torch.full((tg.B, Q), -1, dtype=torch.int32)
        inv = torch.zeros(tg.B, dtype=torch.float32)
        for b, (iq, it) in enumerate(matches):
            m[b, iq] = it.to(torch.int32)
            inv[b] = 1.0 / len(it) if len(it) > 0 else 0.0  # num_masks = target_mask.shape[0], criterion.py:190
        return m, inv

    def loss_labels(self, pred_class, target_classes):
        """criterion.py:141-166"""
        return {"loss_ce": F.cross_entropy(pred_class.float().transpose(1, 2), target_classes, weight=self.empty_weight)}

    def loss_masks(self, logits, tg, stats, match, inv_m):
        """criterion.py:168-213"""
        if self.num_points != -1:  # a fresh random subset for the loss (criterion.py:181-182)
            stats = mask_costs(logits, tg, _row_subset(tg, self.num_points, logits.device))
        lm = _MaskLossFn.apply(logits, tg, stats, match, inv_m)
        return {"loss_mask": lm[0], "loss_dice": lm[1]}

    def _losses_of(self, pred_class, logits, tg, stats, matches):
        Q = pred_class.shape[1]
        dev = pred_class.device
        out = {}
        for name in self.loss_names:
            assert name in ("labels", "masks"), f"Invalid loss '{name}'"
            if name == "labels":
                out.update(self.loss_labels(pred_class, self._target_classes(matches, tg, Q).to(dev, non_blocking=True)))
            else:
                m, inv = self._match_table(matches, tg, Q)
                out.update(self.loss_masks(logits, tg, stats, m.to(dev, non_blocking=True), inv.to(dev, non_blocking=True)))
        return out

    # -- forward -----------------------------------------------------------------------------------------------------
    def forward(self, model_output, targets):
        """criterion.py:237-281.  `model_output`: the reference's ModelOutput, or the dict of mask3d_model.Mask3D."""
        lib.require_cuda()
        cls_all = list(_get(model_output, "pred_class_all_decoders"))
        pm_all = list(_get(model_output, "pred_mask_all_decoders"))
        order = [len(cls_all) - 1] + list(range(len(cls_all) - 1))  # final output first, then mask modules 0..D-2
        sizes = _sizes_of(pm_all[-1], targets)
        dev = cls_all[-1].device
        tg = pack_targets(targets, sizes, dev)
        logits = [_batched_logits(pm, sizes) for pm in pm_all]

        # num_masks (criterion.py:243-252) only feeds arguments the loss functions overwrite or ignore; keep its collective so
        # that ranks running this class and ranks running the reference's stay in step
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            nm = torch.as_tensor([float(sum(tg.n_targets))], dtype=torch.float, device=dev)
            torch.distributed.all_reduce(nm)

        losses = {}
        with lib.stream_scope():
            if self.matcher.num_points == -1:
                with torch.no_grad():
                    stats = {d: mask_costs(logits[d], tg) for d in order}
                    C = torch.stack([self.matcher.cost_matrix(cls_all[d], stats[d], tg) for d in order]).cpu()  # the one sync
                matches = {d: self.matcher.assign(C[i], tg) for i, d in enumerate(order)}
            for i, d in enumerate(order):
                if self.matcher.num_points != -1:  # keep the reference's order of random draws: matcher, then loss, per output
                    m, st = self.matcher.memory_efficient_forward({"pred_class": cls_all[d], "pred_mask": logits[d]}, targets, tg, True)
                    matches, stats = {d: m}, {d: st}
                l_dict = self._losses_of(cls_all[d], logits[d], tg, stats[d], matches[d])
                suffix = "" if i == 0 else f"_mask_module_{d}"
                losses.update({k + suffix: v for k, v in l_dict.items()})
        self.last_matches = matches if self.matcher.num_points == -1 else None
        return losses

    def __repr__(self):
        body = [f"matcher: {self.matcher.__repr__(_repr_indent=8)}", f"losses: {self.loss_names}", f"weight_dict: {self.weight_dict}",
                f"num_classes: {self.num_classes}", f"eos_coef: {self.eos_coef}", f"num_points: {self.num_points}",
                f"oversample_ratio: {self.oversample_ratio}", f"importance_sample_ratio: {self.importance_sample_ratio}"]
        return "\n".join(["Criterion " + self.__class__.__name__] + [" " * 4 + line for line in body])
