"""One 3^3 tensor-core conv launch shape under the ablation / pipeline-shape switches of the environment (read once per process).

usage: python scripts/conv_diag.py <scene> <reps> <cin> [<cout>]
Caches the voxelised scene under /tmp so that a sweep of processes pays the generator once.  Prints one line:
    tag=<MK3D_* settings> cin cout ms TFLOP/s
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from room_instance_segmentation_mask3d_b200 import engine, lib, ops, scenes  # noqa: E402

scene = sys.argv[1]
reps = int(sys.argv[2])
cin = int(sys.argv[3])
cout = int(sys.argv[4]) if len(sys.argv) > 4 else cin
ksize = int(sys.argv[5]) if len(sys.argv) > 5 else 3      # 3 -> 3^3 kernel map, 1 -> 1x1 (identity map)
odt = {"bf16": torch.bfloat16, "f32": torch.float32}[sys.argv[6]] if len(sys.argv) > 6 else torch.bfloat16
use_bias = len(sys.argv) > 7 and sys.argv[7] == "bias"
cache = f"/tmp/mk3d_scene_{scene}.npy"
if os.path.exists(cache):
    c = np.load(cache)
else:
    s = {"s200k": lambda: scenes.box_room(), "floor2": lambda: scenes.floor_plan(1000, voxel_m=0.02),
         "floor4": lambda: scenes.floor_plan(1000, voxel_m=0.04)}[scene]()
    c, _ = scenes.collate([s])
    np.save(cache, c)
c = torch.from_numpy(c).cuda()
lib.load()
if os.environ.get('MK3D_GATHER'):
    lib.call('mk_set_gather_mode', int(os.environ['MK3D_GATHER']))
mgr = engine.CoordinateManager(D=3)
k0, _ = mgr.insert_and_map(c)
km3 = mgr.neighbor_table(k0, k0, ksize, 1)
n0 = mgr.size(k0)
x = torch.randn(n0, cin, device="cuda").bfloat16()
W = torch.randn(ksize ** 3, cin, cout, device="cuda") / (ksize ** 3 * cin) ** 0.5
bias = torch.randn(1, cout, device="cuda") if use_bias else None
fn = lambda: ops.conv_forward(x, W, km3, bias=bias, mode="bf16", out_dtype=odt)
trace = os.environ.get("MK3D_TRACE")
fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    fn()
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / reps
tag = " ".join(f"{k[5:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("MK3D_"))
print(f"tag=[{tag}] K={ksize ** 3} out={odt} bias={use_bias} N={n0} P={km3.num_pairs} {cin}->{cout}: {t:.4f} ms {2 * km3.num_pairs * cin * cout / t / 1e9:.1f} TFLOP/s", flush=True)
