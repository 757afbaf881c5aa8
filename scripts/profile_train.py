"""Kernel-level time breakdown of one Res16UNet34C fwd+bwd step (module path) with torch.profiler; SHARES only."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from room_instance_segmentation_mask3d_b200 import lib, mask3d_model, me_api, ops, scenes, weights  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "s200k"
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16"
path = sys.argv[3] if len(sys.argv) > 3 else "module"
ops.set_compute_mode(mode)
s = scenes.box_room() if scene == "s200k" else scenes.floor_plan(1000, voxel_m=0.02)
c, f = scenes.collate([s])
c, f = torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda()
model = mask3d_model.Res16UNet34C()
weights.load_synth(model)
model = model.cuda().train()


def step():
    for p in model.parameters():
        p.grad = None
    x = me_api.SparseTensor(features=f, coordinates=c)
    out, fm = model.forward_train(x) if path == "fused" else model(x)
    (out.F.float().square().mean() + sum(t.F.float().square().mean() for t in fm[:-1])).backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    if e.device_type == torch.autograd.DeviceType.CUDA or getattr(e, "self_device_time_total", 0) > 0:
        rows.append((e.self_device_time_total, e.count, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"profile_train_{scene}_{mode}_{path}.txt"), "w") as fh:
    hdr = f"scene {scene} voxels {c.shape[0]} mode {mode}: sum of GPU kernel time {tot/1e3:.1f} ms\n"
    fh.write(hdr)
    print(hdr)
    for t, n, k in rows[:45]:
        line = f"{t/1e3:9.3f} ms {100*t/tot:5.1f}%  x{n:<5d} {k[:110]}\n"
        fh.write(line)
        print(line, end="")
