"""A/B of the two A-operand gather mechanisms of k_conv_umma (LDGSTS vs TMA tile::gather4): same inputs, outputs must be bit-identical."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from room_instance_segmentation_mask3d_b200 import engine, lib, ops, scenes  # noqa: E402

lib.load()
ALT = int(sys.argv[1]) if len(sys.argv) > 1 else 2
s = scenes.box_room()
c, _ = scenes.collate([s])
c = torch.from_numpy(c).cuda()
mgr = engine.CoordinateManager(D=3)
k0, _ = mgr.insert_and_map(c)
n0 = mgr.size(k0)
for ks, stride in ((3, 1), (2, 2), (1, 1)):
    if stride == 2:
        k1 = mgr.stride(k0, 2)
        km = mgr.neighbor_table(k0, k1, 2, 1)
    else:
        km = mgr.neighbor_table(k0, k0, ks, 1)
    for cin, cout in ((32, 32), (96, 96), (128, 96), (256, 256), (96, 128)):
        for mode in ("bf16", "tf32"):
            x = torch.randn(n0, cin, device="cuda")
            x = x.bfloat16() if mode == "bf16" else x
            W = torch.randn(ks ** 3, cin, cout, device="cuda") / (ks ** 3 * cin) ** 0.5
            lib.call("mk_set_gather_mode", 0)
            a = ops.conv_forward(x, W, km, mode=mode, out_dtype=torch.float32)
            lib.call("mk_set_gather_mode", ALT)
            b = ops.conv_forward(x, W, km, mode=mode, out_dtype=torch.float32)
            lib.call("mk_set_gather_mode", 0)
            torch.cuda.synchronize()
            print(f"ks {ks} stride {stride} {cin}->{cout} {mode}: max |a-b| = {(a - b).abs().max().item():.3e}  |a| max {a.abs().max().item():.3f}", flush=True)
