"""Mask3D inference at the reference's native scale: batches of B scenes at 15 cm voxels (conf/data/indoor.yaml: voxel_size 0.15,
batch_size 16).  Prints scenes/s and the host enqueue time per batch."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from room_instance_segmentation_mask3d_b200 import lib, mask3d_model, me_api, ops, scenes, weights  # noqa: E402
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
voxel = float(sys.argv[2]) if len(sys.argv) > 2 else 0.15
lib.load(); lib.reserve_device_pool(16.0)
batches = []
for j in range(3):
    c, f = scenes.collate([scenes.floor_plan(1000 + 16 * j + i, voxel_m=voxel) for i in range(B)])
    batches.append((torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda()))
model = mask3d_model.Mask3D().eval(); weights.load_synth(model); model = model.cuda()
for i in range(6):
    model(me_api.SparseTensor(features=batches[i % 3][1], coordinates=batches[i % 3][0]))
torch.cuda.synchronize()
n = 12
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for i in range(n):
    model(me_api.SparseTensor(features=batches[i % 3][1], coordinates=batches[i % 3][0]))
e1.record(); t_enq = (time.perf_counter() - t0) / n
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
nv = sum(b[0].shape[0] for b in batches) / 3
print(f"B={B} voxel={voxel*100:.0f}cm voxels/batch={nv:.0f}: {ms:.2f} ms/batch -> {B / (ms * 1e-3):.0f} scenes/s, {nv / (ms * 1e-3) / 1e6:.1f} M voxels/s; host enqueue {t_enq * 1e3:.2f} ms/batch")
