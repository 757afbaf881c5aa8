"""cProfile of SetCriterion forward + backward at the bench scene's scale (host side of scripts/loss_bench.py)."""
import cProfile
import pstats
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from room_instance_segmentation_mask3d_b200 import criterion as CR  # noqa: E402

n, Q, T, D, C = 889057, 25, 24, 13, 1
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
inst = (torch.arange(n, device=dev) * T // n) % T
masks = torch.stack([inst == t for t in range(T)])
targets = [SimpleNamespace(instances_labels=torch.zeros(T, dtype=torch.long, device=dev), instances_masks=masks)]
pm = [torch.randn(n, Q, device=dev, generator=g).requires_grad_(True) for _ in range(D)]
cls = [torch.randn(1, Q, C + 1, device=dev, generator=g).requires_grad_(True) for _ in range(D)]
out = dict(pred_class=cls[-1], pred_class_all_decoders=cls, pred_mask=pm[-1], pred_mask_all_decoders=pm)
w = {"loss_ce": 2.0, "loss_mask": 5.0, "loss_dice": 2.0}
wd = dict(w)
for d in range(D - 1):
    wd.update({k + f"_mask_module_{d}": v for k, v in w.items()})
crit = CR.SetCriterion(C, CR.HungarianMatcher(2.0, 5.0, 2.0, -1), wd, 0.1, ["labels", "masks"], -1, 3.0, 0.75, -1).cuda()


def step():
    for x in pm + cls:
        x.grad = None
    losses = crit(out, targets)
    sum(losses[k] * wd[k] for k in losses).backward()
    torch.cuda.synchronize()


for _ in range(3):
    step()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
