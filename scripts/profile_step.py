"""Kernel-level time breakdown of one bench step with torch.profiler (CUPTI); writes gpurun_out/profile_step.txt.
Not a bench number: profiler overhead inflates wall time; use it for SHARES only."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from room_instance_segmentation_mask3d_b200 import lib, mask3d_model, me_api, ops, scenes, weights  # noqa: E402

voxel = float(sys.argv[1]) if len(sys.argv) > 1 else 0.02
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16"
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 1  # scenes per step (one batch)
ops.set_compute_mode(mode)
c, f = scenes.collate([scenes.floor_plan(1000 + i, voxel_m=voxel) for i in range(nb)])
c, f = torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda()
model = mask3d_model.Mask3D().eval()
weights.load_synth(model)
model = model.cuda()
for _ in range(3):
    model(me_api.SparseTensor(features=f, coordinates=c))
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    model(me_api.SparseTensor(features=f, coordinates=c))
t_cpu = (time.perf_counter() - t0) / 3
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 3
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    model(me_api.SparseTensor(features=f, coordinates=c))
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    if e.device_type == torch.autograd.DeviceType.CUDA or getattr(e, "self_device_time_total", 0) > 0:
        rows.append((e.self_device_time_total, e.count, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"profile_step_{mode}.txt"), "w") as fh:
    hdr = f"voxels {c.shape[0]} mode {mode}: python enqueue time/step {t_cpu*1e3:.1f} ms, wall/step {t_all*1e3:.1f} ms, sum of GPU kernel time {tot/1e3:.1f} ms\n"
    fh.write(hdr)
    print(hdr)
    for t, n, k in rows[:60]:
        line = f"{t/1e3:9.3f} ms {100*t/tot:5.1f}%  x{n:<5d} {k[:110]}\n"
        fh.write(line)
        print(line, end="")
