"""Mask-logit product microbenchmark: [N0,128] bf16 x [128,25] -> fp32 [N0,25] through k_conv_umma (K = 1, scene-select table)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from room_instance_segmentation_mask3d_b200 import engine, lib, ops, scenes  # noqa: E402
c, _ = scenes.collate([scenes.floor_plan(1000, voxel_m=0.02)])
c = torch.from_numpy(c).cuda()
lib.load()
mgr = engine.CoordinateManager(D=3)
k0, _ = mgr.insert_and_map(c)
sel = mgr.scene_select_table(k0)
n = c.shape[0]
feat = torch.randn(n, 128, device="cuda").bfloat16()
emb = torch.randn(1, 25, 128, device="cuda")
img, _ = ops.mask_embed_image(emb)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ops.mask_logits_from_image(feat, img, sel, 25)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    out = ops.mask_logits_from_image(feat, img, sel, 25)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"mask logits N={n}: {ms*1e3:.1f} us, {(n*128*2 + n*25*4)/ms/1e6:.0f} GB/s (algorithmic bytes in+out)")
