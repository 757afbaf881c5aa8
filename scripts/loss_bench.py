"""Loss side at the bench scene's scale: one 889 k-voxel scene, 25 queries, 24 instances, 13 decoder outputs.

Times (CUDA events, after warm-up) (a) mk_mask_costs and mk_mask_loss_bwd per call against their HBM roofline,
(b) SetCriterion forward + backward (matcher cost matrices for 13 outputs, ONE device->host copy, 13 scipy assignments, losses,
gradients of all 13 logit tensors), (c) the reference's formulation with stock PyTorch ops on the same GPU (fp32 targets,
two BCE maps, sigmoid, three einsums per output, matcher.py:17-56 + criterion.py:27-72) for scale.
"""
import json
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from types import SimpleNamespace  # noqa: E402

from room_instance_segmentation_mask3d_b200 import criterion as CR, lib  # noqa: E402

n, Q, T, D, C = int(sys.argv[1]) if len(sys.argv) > 1 else 889057, 25, 24, 13, 1
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
inst = (torch.arange(n, device=dev) * T // n + (torch.rand(n, device=dev, generator=g) < 0.01).long()) % T
masks = torch.stack([inst == t for t in range(T)])
labels = torch.zeros(T, dtype=torch.long, device=dev)
targets = [SimpleNamespace(instances_labels=labels, instances_masks=masks)]
pm = [(torch.randn(n, Q, device=dev, generator=g) * 3 + (masks[:Q].T.float() * 4 - 2 if T >= Q else 0)).requires_grad_(True) for _ in range(D)]
for x in pm:
    x.data[:, :T] += masks.T.float() * 4 - 2
cls = [torch.randn(1, Q, C + 1, device=dev, generator=g).requires_grad_(True) for _ in range(D)]
out = dict(pred_class=cls[-1], pred_class_all_decoders=cls, pred_mask=pm[-1], pred_mask_all_decoders=pm)
w = {"loss_ce": 2.0, "loss_mask": 5.0, "loss_dice": 2.0}
wd = dict(w)
for d in range(D - 1):
    wd.update({k + f"_mask_module_{d}": v for k, v in w.items()})
matcher = CR.HungarianMatcher(2.0, 5.0, 2.0, -1)
crit = CR.SetCriterion(C, matcher, wd, 0.1, ["labels", "masks"], -1, 3.0, 0.75, -1).cuda()


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters, (time.perf_counter() - t0) * 1e3 / iters


tg = CR.pack_targets(targets, [n], dev)
res = {"voxels": n, "queries": Q, "instances": T, "decoder_outputs": D}
ms, _ = timed(lambda: CR.pack_targets(targets, [n], dev))
res["pack_targets_ms"] = ms
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
i = [0]


def costs():
    i[0] += 1
    CR.mask_costs(pm[i[0] % D], tg)  # a different 89 MB tensor each call: 13 x 89 MB > L2


ms, _ = timed(costs, iters=26)
bytes_costs = n * (4 * Q + 4 * tg.W)
res["mk_mask_costs"] = {"ms": ms, "algorithmic_MB": bytes_costs / 1e6, "GBps": bytes_costs / ms / 1e6, "frac_of_6455": bytes_costs / ms / 1e6 / 6455.3}
st = CR.mask_costs(pm[0], tg)
m, inv = CR.SetCriterion._match_table([(torch.arange(T), torch.arange(T))], tg, Q)
m, inv = m.cuda(), inv.cuda()
coef = torch.rand(1, Q, 3, device=dev)
grads = [torch.empty(n, Q, device=dev) for _ in range(D)]


def bwd():
    i[0] += 1
    k = i[0] % D
    lib.call("mk_mask_loss_bwd", lib.p(pm[k].detach()), n, Q, lib.p(tg.bits), tg.W, None, lib.p(tg.scene_offsets), 1, n, lib.p(m), lib.p(coef),
             lib.p(grads[k]), lib.stream())


ms, _ = timed(bwd, iters=26)
bytes_bwd = n * (8 * Q + 4 * tg.W)
res["mk_mask_loss_bwd"] = {"ms": ms, "algorithmic_MB": bytes_bwd / 1e6, "GBps": bytes_bwd / ms / 1e6, "frac_of_6455": bytes_bwd / ms / 1e6 / 6455.3}


def step():
    for x in pm + cls:
        x.grad = None
    losses = crit(out, targets)
    sum(losses[k] * wd[k] for k in losses).backward()


ms, wall = timed(step, iters=5, warm=2)
res["set_criterion_fwd_bwd"] = {"gpu_ms": ms, "wall_ms": wall}


def stock_step():
    """the reference's formulation (matcher.py:17-56, criterion.py:27-72, 168-213) with stock PyTorch ops on this GPU"""
    for x in pm:
        x.grad = None
    total = 0
    tgt = masks.float()
    for d in range(D):
        with torch.no_grad():
            xm = pm[d].detach().T
            pos = F.binary_cross_entropy_with_logits(xm, torch.ones_like(xm), reduction="none")
            neg = F.binary_cross_entropy_with_logits(xm, torch.zeros_like(xm), reduction="none")
            cm = (torch.einsum("nc,mc->nm", pos, tgt) + torch.einsum("nc,mc->nm", neg, 1 - tgt)) / n
            s = xm.sigmoid()
            cd = 1 - (2 * torch.einsum("nc,mc->nm", s, tgt) + 1) / (s.sum(-1)[:, None] + tgt.sum(-1)[None, :] + 1)
            Cm = (5.0 * cm + 2.0 * cd).cpu()
        from scipy.optimize import linear_sum_assignment
        iq, it = linear_sum_assignment(Cm.numpy())
        x = pm[d][:, torch.as_tensor(iq, device=dev)].T
        tm = tgt[torch.as_tensor(it, device=dev)]
        lm = F.binary_cross_entropy_with_logits(x, tm, reduction="none").mean(1).sum() / len(it)
        sg = x.sigmoid()
        ld = (1 - (2 * (sg * tm).sum(-1) + 1) / (sg.sum(-1) + tm.sum(-1) + 1)).sum() / len(it)
        total = total + 5.0 * lm + 2.0 * ld
    total.backward()


ms, wall = timed(stock_step, iters=3, warm=1)
res["stock_pytorch_same_gpu_fwd_bwd"] = {"gpu_ms": ms, "wall_ms": wall}
print(json.dumps(res))
