"""Whole training step of Mask3D: Mask3D.forward_train (bf16 tensor cores) -> SetCriterion
(matcher costs + Hungarian + losses of all 13 decoder outputs) -> backward -> AdamW (lr 1e-4, the reference's
conf/optimizer/adamw.yaml), on synthetic scenes with synthetic instance targets (runs of voxel rows).

usage: python scripts/train_step_bench.py [scene=floor2|s200k|batch16] [reps]
  floor2   one 889 k-voxel floor at 2 cm          s200k   one 224 k-voxel room at 2 cm
  batch16  16 floors at 15 cm (the reference's native batch: conf/data/indoor.yaml batch_size 16)
Prints one JSON line: ms per phase (CUDA events), voxels/s of the whole step."""
import json
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from room_instance_segmentation_mask3d_b200 import criterion as CR, lib, mask3d_model, me_api, ops, scenes, weights  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "floor2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
if scene == "batch16":
    sc = [scenes.floor_plan(1000 + i, voxel_m=0.15) for i in range(16)]
elif scene == "s200k":
    sc = [scenes.box_room()]
else:
    sc = [scenes.floor_plan(1000, voxel_m=0.02)]
c, f = scenes.collate(sc)
c, f = torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda()
sizes = [len(s) for s in sc]
T = 12
targets = []
for n in sizes:  # T instances per scene = runs of voxel rows (lexicographic order => slabs of the scene)
    seg = torch.div(torch.arange(n, device="cuda") * T, n, rounding_mode="floor")
    targets.append(SimpleNamespace(instances_labels=torch.zeros(T, dtype=torch.long, device="cuda"),
                                   instances_masks=torch.stack([seg == t for t in range(T)])))
lib.load()
lib.reserve_device_pool(64.0)
ops.set_compute_mode("bf16")
model = mask3d_model.Mask3D()
weights.load_synth(model, seed=0)
model = model.cuda().train()
params = [p for p in model.parameters()]
D = model.num_decoders * len(model.hlevels) + 1
w = {"loss_ce": 2.0, "loss_mask": 5.0, "loss_dice": 2.0}
wd = dict(w)
for d in range(D - 1):
    wd.update({k + f"_mask_module_{d}": v for k, v in w.items()})
crit = CR.SetCriterion(model.num_classes, CR.HungarianMatcher(2.0, 5.0, 2.0, -1), wd, 0.1, ["labels", "masks"], -1, 3.0, 0.75, -1).cuda()
opt = torch.optim.AdamW([p for n_, p in model.named_parameters() if not n_.startswith("backbone.final")], lr=1e-4)
ev = lambda: torch.cuda.Event(enable_timing=True)
marks = []


def step(record=False):
    e = [ev() for _ in range(5)] if record else None
    opt.zero_grad(set_to_none=True)
    if record: e[0].record()
    out = model.forward_train(me_api.SparseTensor(features=f, coordinates=c))
    if record: e[1].record()
    losses = crit(out, targets)
    total = sum(losses[k] * wd[k] for k in losses)
    if record: e[2].record()
    total.backward()
    if record: e[3].record()
    opt.step()
    if record:
        e[4].record()
        marks.append(e)
    return float(total.detach()) if record else None


for _ in range(6):
    step()
torch.cuda.synchronize()
vals = [step(True) for _ in range(reps)]
torch.cuda.synchronize()
ph = [sum(m[i].elapsed_time(m[i + 1]) for m in marks) / reps for i in range(4)]
tot = sum(m[0].elapsed_time(m[4]) for m in marks) / reps
n = sum(sizes)
print(json.dumps({"scene": scene, "scenes": len(sizes), "voxels": n, "decoder_outputs": D, "instances_per_scene": T,
                  "ms_forward": ph[0], "ms_criterion": ph[1], "ms_backward": ph[2], "ms_adamw": ph[3], "ms_step": tot,
                  "voxels_per_s": n / (tot * 1e-3), "scenes_per_s": len(sizes) / (tot * 1e-3), "loss_first": vals[0], "loss_last": vals[-1],
                  "what": "Mask3D.forward_train + SetCriterion + backward + AdamW; maps and plans rebuilt every step"}))
