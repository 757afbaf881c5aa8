"""Host-side (Python enqueue) cost of one fused training step of the backbone: cProfile, no synchronize inside."""
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

scene = sys.argv[1] if len(sys.argv) > 1 else "s200k"
s = scenes.box_room() if scene == "s200k" else scenes.floor_plan(1000, voxel_m=0.02)
c, f = scenes.collate([s])
c, f = torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda()
model = mask3d_model.Res16UNet34C()
weights.load_synth(model)
model = model.cuda().train()


def step():
    for p in model.parameters():
        p.grad = None
    out, fm = model.forward_train(me_api.SparseTensor(features=f, coordinates=c))
    (out.F.float().square().mean() + sum(t.F.float().square().mean() for t in fm[:-1])).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
n = 5
t0 = time.perf_counter()
for _ in range(n):
    step()
t_enq = (time.perf_counter() - t0) / n
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / n
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    step()
pr.disable()
torch.cuda.synchronize()
st = io.StringIO()
pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(35)
out = f"{scene} voxels {c.shape[0]}: enqueue {t_enq * 1e3:.2f} ms/step, wall {t_all * 1e3:.2f} ms/step ({n} steps profiled below)\n" + st.getvalue()
open(os.path.join(ROOT, "gpurun_out", "host_profile_train.txt"), "w").write(out)
print(out[:7000])
