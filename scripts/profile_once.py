"""One inference step (and optionally one training step) between cudaProfilerStart / cudaProfilerStop, for
`ncu --profile-from-start off`: every kernel family of the path is launched once in steady state (warm caches, weight images and
CUDA graphs built).

usage: python scripts/profile_once.py [infer|train] [scene=floor2|s200k] [B]"""
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from room_instance_segmentation_mask3d_b200 import criterion as CR  # noqa: E402
from room_instance_segmentation_mask3d_b200 import lib, mask3d_model, me_api, ops, scenes, weights  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "infer"
scene = sys.argv[2] if len(sys.argv) > 2 else "floor2"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ops.set_compute_mode("bf16")
gen = (lambda i: scenes.box_room()) if scene == "s200k" else (lambda i: scenes.floor_plan(1000 + i, voxel_m=0.02))
sc = [gen(i) for i in range(B)]
c, f = scenes.collate(sc)
c, f = torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda()
model = mask3d_model.Mask3D().eval()
weights.load_synth(model)
model = model.cuda()
if what == "infer":
    model.use_cuda_graphs = False  # the query-side launches are then visible to the profiler one by one
    step = lambda: model(me_api.SparseTensor(features=f, coordinates=c), compact=True)
else:
    model.train()
    T = 24
    tgt = []
    for s in sc:
        nb = len(s)
        seg = torch.div(torch.arange(nb, device="cuda") * T, nb, rounding_mode="floor")
        tgt.append(SimpleNamespace(instances_labels=torch.zeros(T, dtype=torch.long, device="cuda"),
                                   instances_masks=torch.stack([seg == t for t in range(T)])))
    wts = {"loss_ce": 2.0, "loss_mask": 5.0, "loss_dice": 2.0}
    wd = dict(wts)
    for d in range(12):
        wd.update({k + f"_mask_module_{d}": v for k, v in wts.items()})
    crit = CR.SetCriterion(1, CR.HungarianMatcher(2.0, 5.0, 2.0, -1), wd, 0.1, ["labels", "masks"], -1, 3.0, 0.75, -1).cuda()

    def step():
        for p in model.parameters():
            p.grad = None
        out = model.forward_train(me_api.SparseTensor(features=f, coordinates=c))
        ls = crit(out, tgt)
        sum(ls[k] * wd[k] for k in ls).backward()

for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one", what, "step:", c.shape[0], "voxels")
