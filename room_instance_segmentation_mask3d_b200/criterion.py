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
    labs = [_get(t, "instances_labels").detach().long() for t in targets]
    flat = (torch.cat(labs) if B else torch.zeros(0, dtype=torch.long)).cpu().numpy()  # ONE device -> host copy for the batch
    labels_host, lab, ign, o = [], np.zeros((B, Tp), dtype=np.int64), np.zeros((B, Tp), dtype=bool), 0
    for b, T in enumerate(n_targets):
        l = flat[o:o + T]
        o += T
        labels_host.append(l)
        ign[b, :T] = l == IGNORE_LABEL
        lab[b, :T] = np.where(l == IGNORE_LABEL, 0, l)
    so = torch.from_numpy(offs.astype(np.int32)).to(device)
    return PackedTargets(bits, so, [int(s) for s in sizes], n_targets, labels_host, torch.from_numpy(lab).to(device),
                         torch.from_numpy(ign).to(device))


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
    """cost_mask / cost_dice / dice_den [..., B, Q, 32 W], n_rows [..., B] (leading dim = decoder outputs when stacked),
    mask_score [B, Q] or None, row_select = the uint8 row mask(s) the sums were taken over (None = all rows)."""
    __slots__ = ("cost_mask", "cost_dice", "dice_den", "mask_score", "n_rows", "row_select")

    def __init__(self, buf, n_rows, mask_score=None, row_select=None):
        self.cost_mask, self.cost_dice, self.dice_den = buf.unbind(-4)
        self.n_rows, self.mask_score, self.row_select = n_rows, mask_score, row_select


def _f32c(logits: torch.Tensor) -> torch.Tensor:
    x = logits.detach()
    return (x if x.dtype == torch.float32 else x.float()).contiguous()


def mask_costs(logits: torch.Tensor, tg: PackedTargets, row_select: Optional[torch.Tensor] = None, want_score: bool = False,
               out: Optional[torch.Tensor] = None, n_rows: Optional[torch.Tensor] = None) -> MaskStats:
    """One pass over fp32 logits [sum N, Q] -> cost_mask / cost_dice / dice_den [B, Q, 32 W], n_rows [B] (+ mask_score [B, Q]).
    `out` [3, B, Q, 32 W] / `n_rows` [B]: write into slices of a caller's stacked buffers."""
    lib.require_cuda()
    x = _f32c(logits)
    n, Q = x.shape
    assert n == tg.bits.shape[0], f"logits have {n} rows, targets were packed for {tg.bits.shape[0]}"
    B, W = tg.B, tg.W
    dev = x.device
    if out is None:
        out = torch.empty((3, B, Q, 32 * W), dtype=torch.float32, device=dev)
    if n_rows is None:
        n_rows = torch.empty((B,), dtype=torch.float32, device=dev)
    assert out.is_contiguous() and tuple(out.shape) == (3, B, Q, 32 * W) and n_rows.is_contiguous()
    score = torch.empty((B, Q), dtype=torch.float32, device=dev) if want_score else None
    ws = lib.scratch(lib.query("mk_mask_costs_workspace_bytes", B, tg.max_rows, Q, W), dev)
    lib.call("mk_mask_costs", lib.p(x), n, Q, lib.p(tg.bits), W, lib.p(row_select), lib.p(tg.scene_offsets), B, tg.max_rows,
             out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), lib.p(score), n_rows.data_ptr(), lib.p(ws), ws.numel(), lib.stream())
    return MaskStats(out, n_rows, score, row_select)


def mask_scores(pred_mask: torch.Tensor) -> torch.Tensor:
    """Per-query mask score of the post-processing (trainer/trainer.py:389-392, before the top-k column permutation):
    sum_n sigmoid(x) [x > 0] / (sum_n [x > 0] + 1e-6) for logits [N, Q] of ONE scene."""
    lib.require_cuda()
    n = int(pred_mask.shape[0])
    dev = pred_mask.device
    tg = PackedTargets(torch.zeros((n, 1), dtype=torch.int32, device=dev), torch.tensor([0, n], dtype=torch.int32, device=dev), [n], [0],
                       [], None, None)
    return mask_costs(pred_mask, tg, want_score=True).mask_score[0]


def get_mask_and_scores(pred_class: torch.Tensor, pred_mask: torch.Tensor, num_queries: int, num_classes: int, topk_per_image: int = -1):
    """Mirror of InstanceSegmentation.get_mask_and_scores (trainer/trainer.py:373-402) for ONE scene.

    pred_class [Q, num_classes] = softmax class probabilities without the no-object column (trainer.py:321-322), pred_mask
    [N, Q] logits.  Returns (score [K], result_pred_mask bool [N, K], classes [K], heatmap fp32 [N, K]) with the reference's
    devices: score on the CPU, the rest where the inputs live.  The per-query reduction over all N voxels — sum(sigmoid(x) *
    (x > 0)) / (sum(x > 0) + 1e-6), the only arithmetic of the function that touches every logit more than once — is
    `mk_mask_costs`' mask score, computed once per query column instead of once per selected (query, class) pair."""
    lib.require_cuda()
    dev = pred_mask.device
    labels = torch.arange(num_classes, device=dev).unsqueeze(0).repeat(num_queries, 1).flatten(0, 1)
    k = topk_per_image if topk_per_image != -1 else num_queries
    scores_per_query, topk_indices = pred_class.flatten(0, 1).topk(k, sorted=True)
    labels_per_query = labels[topk_indices]
    topk_indices = torch.div(topk_indices, num_classes, rounding_mode="floor")
    col_score = mask_scores(pred_mask)  # [Q], one pass over the logits
    sel = pred_mask[:, topk_indices]
    result_pred_mask = sel > 0
    heatmap = sel.float().sigmoid()
    score = scores_per_query.detach().cpu() * col_score[topk_indices].detach().cpu()
    return score, result_pred_mask, labels_per_query, heatmap


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
        """C [..., B, Q, 32 W] on the device (matcher.py:103-160) for pred_class [..., B, Q, C+1]; columns >= T of a scene are
        padding."""
        prob = pred_class.detach().float().softmax(-1)
        lead = (1,) * (prob.dim() - 3)
        shape = (*prob.shape[:-1], tg.labels_dev.shape[1])
        cc = -prob.gather(-1, tg.labels_dev.view(*lead, tg.B, 1, -1).expand(shape))
        cc = torch.where(tg.ignore_dev.view(*lead, tg.B, 1, -1), torch.full_like(cc, -1.0), cc)  # matcher.py:117
        return self.cost_mask * stats.cost_mask + self.cost_class * cc + self.cost_dice * stats.cost_dice

    @staticmethod
    def assign(C_host: torch.Tensor, tg: PackedTargets):
        """Hungarian assignment per scene on the host copy of C [B, Q, Tp] (matcher.py:160-162)."""
        C = C_host.numpy() if isinstance(C_host, torch.Tensor) else C_host
        assert not np.isnan(C).any(), "out mask must not be NaN"  # matcher.py:120
        out = []
        for b in range(tg.B):
            i, j = linear_sum_assignment(C[b, :, : tg.n_targets[b]])
            out.append((torch.from_numpy(i.astype(np.int64)), torch.from_numpy(j.astype(np.int64))))
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
    """(loss_mask, loss_dice) of the matched pairs of D decoder outputs at once, each summed over the scenes of the batch
    (criterion.py:168-213) -> [D, 2].  Forward = gathers from the stacked cost matrices; backward = one mk_mask_loss_bwd per
    output.  `stats` fields carry a leading D dimension, match int32 [D, B, Q], inv_m fp32 [D, B] = 1 / matched pairs."""

    @staticmethod
    def forward(ctx, tg: PackedTargets, stats: MaskStats, match: torch.Tensor, inv_m: torch.Tensor, *logits):
        m = match.long().clamp(min=0).unsqueeze(-1)
        live = ((match >= 0) & (stats.n_rows > 0).unsqueeze(-1)).float()  # no points, no loss (criterion.py:197-203)
        cm = stats.cost_mask.gather(-1, m).squeeze(-1) * live
        cd = stats.cost_dice.gather(-1, m).squeeze(-1) * live
        den = stats.dice_den.gather(-1, m).squeeze(-1)
        ctx.tg = tg
        ctx.row_select = stats.row_select if isinstance(stats.row_select, (list, tuple)) else [stats.row_select] * len(logits)
        ctx.logits = logits
        ctx.save_for_backward(match, inv_m, cd, den, stats.n_rows, live)
        return torch.stack([(cm.sum(-1) * inv_m).sum(-1), (cd.sum(-1) * inv_m).sum(-1)], dim=-1)

    @staticmethod
    def backward(ctx, g):
        match, inv_m, cd, den, n_rows, live = ctx.saved_tensors
        tg = ctx.tg
        g = g.float()
        gm, gd, im = g[:, 0, None, None], g[:, 1, None, None], inv_m.unsqueeze(-1)
        a = gm * im / n_rows.clamp(min=1.0).unsqueeze(-1) * live
        c1 = gd * im * (1.0 - cd) / den * live  # num / den^2 = (1 - cost_dice) / den
        c2 = 2.0 * gd * im / den * live
        coef = torch.stack([a, c1, c2], dim=-1).contiguous()  # [D, B, Q, 3]
        match = match.contiguous()
        grads = []
        st = lib.stream()
        for d, lg in enumerate(ctx.logits):
            if not ctx.needs_input_grad[4 + d]:
                grads.append(None)
                continue
            x = _f32c(lg)
            grad = torch.empty_like(x)
            n, Q = x.shape
            lib.call("mk_mask_loss_bwd", lib.p(x), n, Q, lib.p(tg.bits), tg.W, lib.p(ctx.row_select[d]), lib.p(tg.scene_offsets), tg.B,
                     tg.max_rows, match[d].data_ptr(), coef[d].data_ptr(), lib.p(grad), st)
            grads.append(grad if lg.dtype == torch.float32 else grad.to(lg.dtype))
        return (None, None, None, None, *grads)


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

    # -- pieces, each for D decoder outputs at once (host work per step, not per output) -------------------------------------
    def _tables(self, matches_all, tg: PackedTargets, Q: int):
        """matches_all[d][b] = (query idx, target idx) -> target classes int64 [D, B, Q] (criterion.py:152-158), matched target
        int32 [D, B, Q] (-1 = none), 1 / matched pairs fp32 [D, B] (num_masks = target_mask.shape[0], criterion.py:190)."""
        D = len(matches_all)
        tc = np.full((D, tg.B, Q), self.num_classes, dtype=np.int64)
        m = np.full((D, tg.B, Q), -1, dtype=np.int32)
        inv = np.zeros((D, tg.B), dtype=np.float32)
        for d, matches in enumerate(matches_all):
            for b, (iq, it) in enumerate(matches):
                iq, it = iq.numpy(), it.numpy()
                tc[d, b, iq] = tg.labels_host[b][it]
                m[d, b, iq] = it
                inv[d, b] = 1.0 / len(it) if len(it) > 0 else 0.0
        return torch.from_numpy(tc), torch.from_numpy(m), torch.from_numpy(inv)

    @staticmethod
    def _match_table(matches, tg: PackedTargets, Q: int):
        """One decoder output: matched target int32 [B, Q] and 1 / matched pairs [B]."""
        m = torch.full((tg.B, Q), -1, dtype=torch.int32)
        inv = torch.zeros(tg.B, dtype=torch.float32)
        for b, (iq, it) in enumerate(matches):
            m[b, iq] = it.to(torch.int32)
            inv[b] = 1.0 / len(it) if len(it) > 0 else 0.0
        return m, inv

    def loss_labels(self, pred_class: torch.Tensor, target_classes: torch.Tensor) -> torch.Tensor:
        """criterion.py:141-166 for pred_class [D, B, Q, C+1], target_classes [D, B, Q] -> [D]: F.cross_entropy with class
        weights and mean reduction = sum(w[target] nll) / sum(w[target]), per decoder output."""
        D = pred_class.shape[0]
        ce = F.cross_entropy(pred_class.float().flatten(0, 1).transpose(1, 2), target_classes.flatten(0, 1), weight=self.empty_weight,
                             reduction="none")
        return ce.view(D, -1).sum(1) / self.empty_weight[target_classes].view(D, -1).sum(1)

    def loss_masks(self, logits: Sequence[torch.Tensor], tg, stats: MaskStats, match, inv_m) -> torch.Tensor:
        """criterion.py:168-213 for D decoder outputs -> [D, 2] = (loss_mask, loss_dice)."""
        if self.num_points != -1:  # a fresh random subset for the loss (criterion.py:181-182), drawn per output in order
            D, dev = len(logits), logits[0].device
            buf = torch.empty((D, 3, tg.B, logits[0].shape[1], 32 * tg.W), dtype=torch.float32, device=dev)
            nr = torch.empty((D, tg.B), dtype=torch.float32, device=dev)
            sels = [_row_subset(tg, self.num_points, dev) for _ in range(D)]
            for d in range(D):
                mask_costs(logits[d], tg, sels[d], out=buf[d], n_rows=nr[d])
            stats = MaskStats(buf, nr, None, sels)
        else:
            rs = stats.row_select if isinstance(stats.row_select, (list, tuple)) else [stats.row_select]
            if any(r is not None for r in rs):
                # the matcher sub-sampled rows (matcher.num_points != -1) but the criterion must not (num_points == -1): the loss
                # is taken over ALL points, as in the reference (criterion.py:181-196), so the statistics are recomputed on all rows
                D, dev = len(logits), logits[0].device
                buf = torch.empty((D, 3, tg.B, logits[0].shape[1], 32 * tg.W), dtype=torch.float32, device=dev)
                nr = torch.empty((D, tg.B), dtype=torch.float32, device=dev)
                for d in range(D):
                    mask_costs(logits[d], tg, None, out=buf[d], n_rows=nr[d])
                stats = MaskStats(buf, nr)
        # NB: a scene without targets contributes 0 to the mask / dice losses here; the reference divides 0 by 0 there (NaN)
        return _MaskLossFn.apply(tg, stats, match, inv_m, *logits)

    def _losses(self, cls_stack, logits, tg, stats, matches_all):
        Q = cls_stack.shape[2]
        dev = cls_stack.device
        tc, m, inv = self._tables(matches_all, tg, Q)
        out = {}
        for name in self.loss_names:
            assert name in ("labels", "masks"), f"Invalid loss '{name}'"
            if name == "labels":
                out["loss_ce"] = self.loss_labels(cls_stack, tc.to(dev, non_blocking=True))
            else:
                lm = self.loss_masks(logits, tg, stats, m.to(dev, non_blocking=True), inv.to(dev, non_blocking=True))
                out["loss_mask"], out["loss_dice"] = lm[:, 0], lm[:, 1]
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
        logits = [_batched_logits(pm_all[d], sizes) for d in order]
        cls = [cls_all[d] for d in order]
        D, Q = len(order), cls[0].shape[1]

        # num_masks (criterion.py:243-252) only feeds arguments the loss functions overwrite or ignore; keep its collective so
        # that ranks running this class and ranks running the reference's stay in step
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            nm = torch.as_tensor([float(sum(tg.n_targets))], dtype=torch.float, device=dev)
            torch.distributed.all_reduce(nm)

        per_output = {}
        with lib.stream_scope():
            if self.matcher.num_points == -1:
                with torch.no_grad():
                    buf = torch.empty((D, 3, tg.B, Q, 32 * tg.W), dtype=torch.float32, device=dev)
                    nr = torch.empty((D, tg.B), dtype=torch.float32, device=dev)
                    for i in range(D):
                        mask_costs(logits[i], tg, out=buf[i], n_rows=nr[i])
                    stats = MaskStats(buf, nr)
                    C = self.matcher.cost_matrix(torch.stack(cls), stats, tg).cpu().numpy()  # the one device -> host copy of the step
                matches_all = [self.matcher.assign(C[i], tg) for i in range(D)]
                per_output = self._losses(torch.stack(cls), logits, tg, stats, matches_all)
            else:  # keep the reference's order of random draws: matcher subset, then loss subset, output by output
                matches_all, rows = [], []
                for i in range(D):
                    m, st = self.matcher.memory_efficient_forward({"pred_class": cls[i], "pred_mask": logits[i]}, targets, tg, True)
                    matches_all.append(m)
                    one = self._losses(cls[i][None], [logits[i]], tg, MaskStats(torch.stack([st.cost_mask, st.cost_dice, st.dice_den])[None],
                                                                             st.n_rows[None], None, [st.row_select]), [m])
                    rows.append(one)
                per_output = {k: torch.cat([r[k] for r in rows]) for k in rows[0]}
        losses = {}
        for i, d in enumerate(order):
            suffix = "" if i == 0 else f"_mask_module_{d}"
            for k, v in per_output.items():
                losses[k + suffix] = v[i]
        self.last_matches = {d: matches_all[i] for i, d in enumerate(order)}
        return losses

    def __repr__(self):
        body = [f"matcher: {self.matcher.__repr__(_repr_indent=8)}", f"losses: {self.loss_names}", f"weight_dict: {self.weight_dict}",
                f"num_classes: {self.num_classes}", f"eos_coef: {self.eos_coef}", f"num_points: {self.num_points}",
                f"oversample_ratio: {self.oversample_ratio}", f"importance_sample_ratio: {self.importance_sample_ratio}"]
        return "\n".join(["Criterion " + self.__class__.__name__] + [" " * 4 + line for line in body])
