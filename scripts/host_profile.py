"""Host-side (Python enqueue) cost of one bench step: cProfile over the fused Mask3D forward, GPU idle-wait excluded by
measuring enqueue time only (no synchronize inside the profiled region).  Writes gpurun_out/host_profile.txt."""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from room_instance_segmentation_mask3d_b200 import lib, mask3d_model, me_api, ops, scenes, weights  # noqa: E402

voxel = float(sys.argv[1]) if len(sys.argv) > 1 else 0.02
c, f = scenes.collate([scenes.floor_plan(1000, voxel_m=voxel)])
c, f = torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda()
model = mask3d_model.Mask3D().eval()
weights.load_synth(model)
model = model.cuda()
for _ in range(3):
    model(me_api.SparseTensor(features=f, coordinates=c))
torch.cuda.synchronize()
n = 5
t0 = time.perf_counter()
for _ in range(n):
    model(me_api.SparseTensor(features=f, coordinates=c))
t_enq = (time.perf_counter() - t0) / n
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / n
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    model(me_api.SparseTensor(features=f, coordinates=c))
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
out = f"voxels {c.shape[0]}: enqueue {t_enq * 1e3:.2f} ms/step, wall {t_all * 1e3:.2f} ms/step ({n} steps profiled below)\n" + s.getvalue()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "host_profile.txt"), "w").write(out)
print(out[:6000])
