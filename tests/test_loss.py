"""Loss side of the mask logits.

CPU: oracle/loss_oracle.py replays tests/golden/loss_tiny.npz, which the reference's UNMODIFIED HungarianMatcher +
SetCriterion wrote (tests/golden/make_golden_loss.py).  GPU: the CUDA path (criterion.py -> mk_target_bits / mk_mask_costs /
mk_mask_loss_bwd through the C ABI) against the same fixture and against the oracle on ragged / edge-case inputs.

Tolerances: floating point, BASELINE.json north_star bar of 1e-3 relative (we assert 2e-4 relative + 1e-6 absolute on costs
and losses, 1e-3 relative + 1e-7 absolute on gradients); assignments exact.
"""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import loss_oracle as LO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_tiny.npz")
W_CLASS, W_MASK, W_DICE, EOS = 2.0, 5.0, 2.0, 0.1  # conf/matcher/hungarian_matcher.yaml, conf/loss/set_criterion.yaml


def _load():
    fx = np.load(GOLD)
    B, D, C = int(fx["B"]), int(fx["D"]), int(fx["num_classes"])
    masks = [torch.from_numpy(fx[f"masks_{b}"]) for b in range(B)]
    labels = [torch.from_numpy(fx[f"labels_{b}"]) for b in range(B)]
    cls = [torch.from_numpy(fx[f"pred_class_{d}"]) for d in range(D)]
    pm = [[torch.from_numpy(fx[f"pred_mask_{d}_{b}"]) for b in range(B)] for d in range(D)]
    return fx, B, D, C, masks, labels, cls, pm


def _weights(D):
    w = {"loss_ce": W_CLASS, "loss_mask": W_MASK, "loss_dice": W_DICE}
    out = dict(w)
    for d in range(D - 1):
        out.update({k + f"_mask_module_{d}": v for k, v in w.items()})
    return out


# ------------------------------------------------------------------------------------------------------------------
# CPU: the oracle is pinned by the reference's own output
# ------------------------------------------------------------------------------------------------------------------
def test_loss_oracle_replays_reference_golden():
    fx, B, D, C, masks, labels, cls, pm = _load()
    for d in range(D):
        for b in range(B):
            c = LO.matcher_costs(cls[d][b], pm[d][b], labels[b], masks[b])
            np.testing.assert_allclose(c["cost_mask"].numpy(), fx[f"cost_mask_{d}_{b}"], rtol=2e-5, atol=1e-6)
            np.testing.assert_allclose(c["cost_dice"].numpy(), fx[f"cost_dice_{d}_{b}"], rtol=2e-5, atol=1e-6)
    for p_ in cls:
        p_.requires_grad_(True)
    for row in pm:
        for x in row:
            x.requires_grad_(True)
    losses, matches = LO.set_criterion(cls, pm, labels, masks, C, EOS, W_CLASS, W_MASK, W_DICE)
    for d in range(D):
        for b in range(B):
            assert np.array_equal(matches[d][b][0], fx[f"match_q_{d}_{b}"]) and np.array_equal(matches[d][b][1], fx[f"match_t_{d}_{b}"])
    gold = dict(zip(fx["loss_names"].tolist(), fx["loss_values"].tolist()))
    assert sorted(losses) == sorted(gold)
    for k, v in gold.items():
        assert abs(float(losses[k]) - v) <= 2e-5 * abs(v) + 1e-6, k
    wd = _weights(D)
    total = sum(losses[k] * wd[k] for k in losses)
    total.backward()
    assert abs(float(total) - float(fx["total"])) <= 2e-5 * abs(float(fx["total"]))
    for d in range(D):
        np.testing.assert_allclose(cls[d].grad.numpy(), fx[f"grad_class_{d}"], rtol=1e-4, atol=1e-7)
        for b in range(B):
            np.testing.assert_allclose(pm[d][b].grad.numpy(), fx[f"grad_mask_{d}_{b}"], rtol=1e-4, atol=1e-8)


def test_loss_oracle_matcher_ignore_label_and_empty_targets():
    fx = np.load(GOLD)
    pc = torch.from_numpy(fx["case2_pred_class"])
    for b in range(3):
        masks, labels = torch.from_numpy(fx[f"case2_masks_{b}"]), torch.from_numpy(fx[f"case2_labels_{b}"])
        c = LO.matcher_costs(pc[b], torch.from_numpy(fx[f"case2_pred_mask_{b}"]), labels, masks)
        C = W_MASK * c["cost_mask"] + W_CLASS * c["cost_class"] + W_DICE * c["cost_dice"]
        i, j = LO.hungarian(C.reshape(pc.shape[1], -1))
        assert np.array_equal(i, fx[f"case2_match_q_{b}"]) and np.array_equal(j, fx[f"case2_match_t_{b}"])


def test_criterion_mirror_has_reference_signature_and_fails_without_cuda():
    import inspect

    from room_instance_segmentation_mask3d_b200 import criterion as CR

    assert list(inspect.signature(CR.HungarianMatcher.__init__).parameters)[1:] == ["cost_class", "cost_mask", "cost_dice", "num_points"]
    assert list(inspect.signature(CR.SetCriterion.__init__).parameters)[1:] == [
        "num_classes", "matcher", "weight_dict", "eos_coef", "loss_names", "num_points", "oversample_ratio", "importance_sample_ratio",
        "class_weights"]  # models/criterion.py:100-111
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            CR.mask_scores(torch.zeros(4, 3))


def test_criterion_host_logic_batched_logits_and_match_tables():
    """Host side of the mirror (no kernel involved): per-scene logit lists that are the row blocks of one tensor are used
    without a copy, anything else is concatenated; the match / class tables follow criterion.py:152-158,190."""
    from room_instance_segmentation_mask3d_b200 import criterion as CR

    base = torch.randn(10, 4, requires_grad=True)
    views = list(base.split([3, 7]))
    assert CR._batched_logits(views, [3, 7]) is base
    assert CR._batched_logits(base, [3, 7]) is base
    sep = [torch.randn(3, 4), torch.randn(7, 4)]
    cat = CR._batched_logits(sep, [3, 7])
    assert cat.shape == (10, 4) and torch.equal(cat[3:], sep[1])
    assert CR._batched_logits(list(base[:, :2].split([3, 7])), [3, 7]).shape == (10, 2)  # column slices: not row blocks, copied
    assert CR._sizes_of(sep, None) == [3, 7]

    tg = SimpleNamespace(B=2, labels_host=[np.array([5, 6, 7]), np.array([1])])
    crit = CR.SetCriterion(9, CR.HungarianMatcher(1, 1, 1, -1), {}, 0.1, ["labels", "masks"], -1, 3.0, 0.75, -1)
    matches = [[(torch.tensor([4, 0, 2]), torch.tensor([0, 2, 1])), (torch.tensor([3]), torch.tensor([0]))]]
    tc, m, inv = crit._tables(matches, tg, 5)
    assert tc.tolist() == [[[7, 9, 6, 9, 5], [9, 9, 9, 1, 9]]]  # unmatched queries -> class num_classes
    assert m.tolist() == [[[2, -1, 1, -1, 0], [-1, -1, -1, 0, -1]]]
    assert torch.allclose(inv, torch.tensor([[1 / 3, 1.0]]))
    ce = crit.loss_labels(torch.zeros(1, 2, 5, 10), tc)  # uniform logits: weighted mean of log(10)
    assert abs(float(ce[0]) - np.log(10)) < 1e-5


# ------------------------------------------------------------------------------------------------------------------
# GPU: the CUDA path
# ------------------------------------------------------------------------------------------------------------------
def _criterion(C, D, num_points=-1):
    from room_instance_segmentation_mask3d_b200 import criterion as CR

    m = CR.HungarianMatcher(cost_class=W_CLASS, cost_mask=W_MASK, cost_dice=W_DICE, num_points=num_points)
    return CR.SetCriterion(num_classes=C, matcher=m, weight_dict=_weights(D), eos_coef=EOS, loss_names=["labels", "masks"],
                           num_points=num_points, oversample_ratio=3.0, importance_sample_ratio=0.75, class_weights=-1).cuda()


def _targets(masks, labels, dev="cuda"):
    return [SimpleNamespace(instances_labels=l.to(dev), instances_masks=m.to(dev)) for m, l in zip(masks, labels)]


@pytest.mark.gpu
@pytest.mark.parametrize("batched", [False, True])
def test_set_criterion_matches_reference_golden(batched):
    fx, B, D, C, masks, labels, cls, pm = _load()
    cls = [c.cuda().requires_grad_(True) for c in cls]
    if batched:  # what mask3d_model.Mask3D returns: one [sum N, Q] tensor per decoder output
        pm = [torch.cat(row).cuda().requires_grad_(True) for row in pm]
    else:  # what the reference's Mask3D returns: per-scene tensors
        pm = [[x.cuda().requires_grad_(True) for x in row] for row in pm]
    crit = _criterion(C, D)
    out = dict(pred_class=cls[-1], pred_class_all_decoders=cls, pred_mask=pm[-1], pred_mask_all_decoders=pm)
    losses = crit(out if batched else SimpleNamespace(**out), _targets(masks, labels))
    gold = dict(zip(fx["loss_names"].tolist(), fx["loss_values"].tolist()))
    assert sorted(losses) == sorted(gold)
    for k, v in gold.items():
        assert abs(float(losses[k]) - v) <= 2e-4 * abs(v) + 1e-6, (k, float(losses[k]), v)
    for d in range(D):
        for b in range(B):
            iq, it = crit.last_matches[d][b]
            assert np.array_equal(iq.numpy(), fx[f"match_q_{d}_{b}"]) and np.array_equal(it.numpy(), fx[f"match_t_{d}_{b}"])
    wd = _weights(D)
    sum(losses[k] * wd[k] for k in losses).backward()
    sizes = [m.shape[1] for m in masks]
    for d in range(D):
        np.testing.assert_allclose(cls[d].grad.cpu().numpy(), fx[f"grad_class_{d}"], rtol=1e-3, atol=1e-7)
        g = pm[d].grad.cpu().split(sizes) if batched else [x.grad.cpu() for x in pm[d]]
        for b in range(B):
            np.testing.assert_allclose(g[b].numpy(), fx[f"grad_mask_{d}_{b}"], rtol=1e-3, atol=1e-7)


@pytest.mark.gpu
def test_matcher_ignore_label_and_scene_without_targets():
    from room_instance_segmentation_mask3d_b200 import criterion as CR

    fx = np.load(GOLD)
    masks = [torch.from_numpy(fx[f"case2_masks_{b}"]) for b in range(3)]
    labels = [torch.from_numpy(fx[f"case2_labels_{b}"]) for b in range(3)]
    pm = [torch.from_numpy(fx[f"case2_pred_mask_{b}"]).cuda() for b in range(3)]
    m = CR.HungarianMatcher(cost_class=W_CLASS, cost_mask=W_MASK, cost_dice=W_DICE, num_points=-1)
    got = m(SimpleNamespace(pred_class=torch.from_numpy(fx["case2_pred_class"]).cuda(), pred_mask=pm), _targets(masks, labels))
    for b in range(3):
        assert np.array_equal(got[b][0].numpy(), fx[f"case2_match_q_{b}"]) and np.array_equal(got[b][1].numpy(), fx[f"case2_match_t_{b}"])


def _stack1(CR, st):
    """MaskStats of one decoder output -> the stacked (D = 1) form _MaskLossFn takes."""
    return CR.MaskStats(torch.stack([st.cost_mask, st.cost_dice, st.dice_den])[None], st.n_rows[None], None, [st.row_select])


def _random_case(seed, sizes, n_targets, Q, overlap=True):
    g = torch.Generator().manual_seed(seed)
    masks, pm = [], []
    for n, T in zip(sizes, n_targets):
        if T == 0:
            masks.append(torch.zeros((0, n), dtype=torch.bool))
        else:
            inst = torch.randint(0, T, (max(n // 13, 1),), generator=g).repeat_interleave(13)[:n]
            inst = torch.cat([inst, torch.zeros(n - len(inst), dtype=torch.long)])
            m = torch.stack([inst == t for t in range(T)])
            if overlap and n > 0:
                m |= torch.rand(T, n, generator=g) < 0.02
            masks.append(m)
        pm.append(torch.randn(n, Q, generator=g) * 3 + (torch.rand(1, Q, generator=g) - 0.5) * 8)
    return masks, pm


@pytest.mark.gpu
@pytest.mark.parametrize("sizes,n_targets,Q", [
    ((5000, 1, 2049, 0, 333), (7, 1, 33, 2, 0), 25),   # ragged: one-voxel scene, empty scene, T > 32 (two words), no targets
    ((4097, 2048), (64, 40), 100),                     # Q > 32 (4 query chunks), T = 64
    ((1500,), (3,), 1),
])
def test_mask_costs_and_scores_vs_oracle(sizes, n_targets, Q):
    from room_instance_segmentation_mask3d_b200 import criterion as CR

    masks, pm = _random_case(3, sizes, n_targets, Q)
    tg = CR.pack_targets([SimpleNamespace(instances_labels=torch.zeros(T, dtype=torch.long), instances_masks=m) for m, T in zip(masks, n_targets)],
                         sizes, "cuda")
    st = CR.mask_costs(torch.cat(pm).cuda(), tg, want_score=True)
    for b, (n, T) in enumerate(zip(sizes, n_targets)):
        if n == 0:
            continue
        c = LO.matcher_costs(torch.zeros(Q, 2), pm[b], torch.zeros(T, dtype=torch.long), masks[b])
        np.testing.assert_allclose(st.cost_mask[b, :, :T].cpu().numpy(), c["cost_mask"].numpy(), rtol=2e-4, atol=1e-6)
        np.testing.assert_allclose(st.cost_dice[b, :, :T].cpu().numpy(), c["cost_dice"].numpy(), rtol=2e-4, atol=1e-6)
        np.testing.assert_allclose(st.mask_score[b].cpu().numpy(), LO.mask_scores(pm[b]).numpy(), rtol=2e-4, atol=1e-6)
        assert float(st.n_rows[b]) == n
        np.testing.assert_allclose(CR.mask_scores(pm[b].cuda()).cpu().numpy(), LO.mask_scores(pm[b]).numpy(), rtol=2e-4, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("num_classes,topk", [(1, -1), (3, -1), (3, 40)])
def test_get_mask_and_scores_vs_oracle(num_classes, topk):
    """Post-processing mask scoring (trainer/trainer.py:373-402): scores, selected (query, class) pairs, binary masks, heatmap."""
    from room_instance_segmentation_mask3d_b200 import criterion as CR

    g = torch.Generator().manual_seed(21)
    n, Q = 7001, 25
    pm = torch.randn(n, Q, generator=g) * 3
    pm[:, 4] = -5.0  # a query whose mask is empty: score 0 / 1e-6
    pc = torch.rand(Q, num_classes + 1, generator=g).softmax(-1)[:, :-1]
    ws, wm, wc, wh = LO.get_mask_and_scores(pc, pm, Q, num_classes, topk)
    gs, gm, gc, gh = CR.get_mask_and_scores(pc.cuda(), pm.cuda(), Q, num_classes, topk)
    assert not gs.is_cuda and torch.equal(gc.cpu(), wc) and torch.equal(gm.cpu(), wm)
    np.testing.assert_allclose(gs.numpy(), ws.numpy(), rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(gh.cpu().numpy(), wh.numpy(), rtol=1e-5, atol=1e-7)


@pytest.mark.gpu
def test_mask_losses_forward_backward_vs_oracle_with_row_subset():
    """Matched losses + gradient on a ragged batch, all points and an explicit row subset (the num_points != -1 path)."""
    from room_instance_segmentation_mask3d_b200 import criterion as CR

    sizes, n_targets, Q = (3001, 700, 64), (6, 35, 0), 25
    masks, pm = _random_case(5, sizes, n_targets, Q)
    g = torch.Generator().manual_seed(9)
    matches = [(torch.randperm(Q, generator=g)[: min(T, Q)].sort().values, torch.randperm(T, generator=g)[: min(T, Q)]) for T in n_targets]
    tg = CR.pack_targets([SimpleNamespace(instances_labels=torch.zeros(T, dtype=torch.long), instances_masks=m) for m, T in zip(masks, n_targets)],
                         sizes, "cuda")
    for subset in (False, True):
        rows = [torch.randperm(n, generator=g)[: n // 2].sort().values if subset else None for n in sizes]
        sel = None
        if subset:
            sel = torch.zeros(sum(sizes), dtype=torch.uint8)
            off = 0
            for n, r in zip(sizes, rows):
                sel[r + off] = 1
                off += n
            sel = sel.cuda()
        x = torch.cat(pm).cuda().requires_grad_(True)
        st = CR.mask_costs(x, tg, sel)
        m, inv = CR.SetCriterion._match_table(matches, tg, Q)
        lm = CR._MaskLossFn.apply(tg, _stack1(CR, st), m.cuda()[None], inv.cuda()[None], x)[0]
        (3.0 * lm[0] + 0.7 * lm[1]).backward()
        ref_x = [p_.clone().requires_grad_(True) for p_ in pm]
        ref = [LO.mask_losses(ref_x[b], masks[b], *matches[b], rows=rows[b]) for b in range(len(sizes))]
        rm, rd = sum(a for a, _ in ref), sum(a for _, a in ref)
        (3.0 * rm + 0.7 * rd).backward()
        assert abs(float(lm[0]) - float(rm)) <= 2e-4 * abs(float(rm)) + 1e-6
        assert abs(float(lm[1]) - float(rd)) <= 2e-4 * abs(float(rd)) + 1e-6
        got = x.grad.cpu().split(list(sizes))
        for b in range(len(sizes)):
            want = ref_x[b].grad if ref_x[b].grad is not None else torch.zeros_like(ref_x[b])
            np.testing.assert_allclose(got[b].numpy(), want.numpy(), rtol=1e-3, atol=1e-8)


@pytest.mark.gpu
def test_set_criterion_random_subset_runs_and_is_seed_deterministic():
    """num_points != -1: the per-item randperm subsets (matcher.py:125-127, criterion.py:181-182) are drawn from the device
    generator in the reference's order; same seed -> same losses, and they stay close to the all-points losses."""
    fx, B, D, C, masks, labels, cls, pm = _load()
    cls = [c.cuda() for c in cls]
    pm = [[x.cuda() for x in row] for row in pm]
    out = SimpleNamespace(pred_class=cls[-1], pred_class_all_decoders=cls, pred_mask=pm[-1], pred_mask_all_decoders=pm)
    crit = _criterion(C, D, num_points=0.5)
    res = []
    for _ in range(2):
        torch.manual_seed(123)
        res.append({k: float(v) for k, v in crit(out, _targets(masks, labels)).items()})
    assert res[0] == res[1]
    gold = dict(zip(fx["loss_names"].tolist(), fx["loss_values"].tolist()))
    for k in ("loss_mask", "loss_dice"):
        assert abs(res[0][k] - gold[k]) < 0.25 * abs(gold[k]) + 0.05


@pytest.mark.gpu
def test_mask_costs_full_size_properties():
    """889 k voxels x 25 queries x 24 instances (the bench scene's scale), where the oracle would take minutes:
    size-independent properties.  (i) a partition of the scene into instances: sum_t sum_{n in t} x = sum_n x, i.e.
    N * sum_t (cost_mask[q,t] - cost_mask[q,empty]) = -sum_n x[n,q];  (ii) the result does not depend on how the rows are cut
    into scenes (one scene vs. the same rows as 3 scenes summed);  (iii) gradient of a linear functional equals finite
    differences of the forward along a random direction."""
    from room_instance_segmentation_mask3d_b200 import criterion as CR

    n, Q, T = 889057, 25, 24
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(n, Q, device="cuda", generator=g) * 3 + 1.0  # mean 1: the column sums of (i) are ~n, far above rounding
    inst = (torch.arange(n, device="cuda") * T // n + (torch.rand(n, device="cuda", generator=g) < 0.01).long()) % T
    masks = torch.stack([inst == t for t in range(T)])
    mk = lambda m: SimpleNamespace(instances_labels=torch.zeros(m.shape[0], dtype=torch.long), instances_masks=m)
    tg1 = CR.pack_targets([mk(masks)], [n], "cuda")
    s1 = CR.mask_costs(x, tg1)
    cm = s1.cost_mask[0].double()  # [Q, 32]: columns >= T are the empty target
    lhs = n * (cm[:, :T] - cm[:, T:T + 1]).sum(1)
    torch.testing.assert_close(lhs, -x.double().sum(0), rtol=1e-4, atol=5.0)
    cuts = [300000, 289057, 300000]
    tg3 = CR.pack_targets([mk(m) for m in masks.split(cuts, dim=1)], cuts, "cuda")
    s3 = CR.mask_costs(x, tg3)
    w = torch.tensor(cuts, device="cuda", dtype=torch.float64)[:, None, None] / n
    torch.testing.assert_close((s3.cost_mask.double() * w).sum(0), s1.cost_mask[0].double(), rtol=1e-5, atol=1e-7)
    # (iii)
    matches = [(torch.arange(T), torch.arange(T))]
    m, inv = CR.SetCriterion._match_table(matches, tg1, Q)
    xr = x.clone().requires_grad_(True)
    lm = CR._MaskLossFn.apply(tg1, _stack1(CR, CR.mask_costs(xr, tg1)), m.cuda()[None], inv.cuda()[None], xr)[0]
    (lm[0] + lm[1]).backward()
    d = xr.grad.sign()  # aligned direction: the directional derivative is sum |grad| ~ 0.5, far above fp32 noise of the loss
    eps = 1e-2
    f = lambda xx: float(CR._MaskLossFn.apply(tg1, _stack1(CR, CR.mask_costs(xx, tg1)), m.cuda()[None], inv.cuda()[None], xx).double().sum())
    fd = (f(x + eps * d) - f(x - eps * d)) / (2 * eps)
    an = float((xr.grad.double() * d.double()).sum())
    assert abs(fd - an) <= 2e-2 * abs(an) + 1e-5, (fd, an)
