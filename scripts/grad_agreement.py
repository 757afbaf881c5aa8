"""Gradient agreement of the bf16 paths with the fp32 module path on one scene (kernel-gradient cosines)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from room_instance_segmentation_mask3d_b200 import mask3d_model as MM, me_api as ME, scenes, weights  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
if which == "tiny":
    sc = [scenes.tiny_scene(23), scenes.tiny_scene(24, (20, 16, 10)), scenes.tiny_scene(25, (9, 30, 7), 1)]
else:
    sc = [scenes.floor_plan(1000, voxel_m=0.06)]
c, f = scenes.collate(sc)
coords, feats = torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda()
model = MM.Res16UNet34C(); weights.load_synth(model, seed=2); model = model.cuda().train()


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


def run(path, mode):
    ME.set_compute_mode(mode)
    for p in model.parameters():
        p.grad = None
    out, fm = path(ME.SparseTensor(feats, coordinates=coords))
    ts = [out] + fm[:-1]
    tgt = [torch.randn(t.F.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7 + i)) for i, t in enumerate(ts)]
    sum((t.F.float() - 0.5 * d).square().mean() for t, d in zip(ts, tgt)).backward()
    return out.F.detach().float(), {n: p.grad.detach().clone().float() for n, p in model.named_parameters() if p.grad is not None}


y0, g0 = run(model.forward, "fp32")
print("voxels", coords.shape[0])
for name, path, mode in (("module bf16", model.forward, "bf16"), ("fused bf16", model.forward_train, "bf16")):
    y, g = run(path, mode)
    ks = sorted((cos(g[n], g0[n]), n) for n in g0 if n.endswith("kernel"))
    ns = [cos(g[n], g0[n]) for n in g0 if not n.endswith("kernel")]
    print(f"{name}: out rel err {float((y - y0).abs().max() / y0.abs().max()):.3e}; kernel-grad cos min {ks[0][0]:.4f} ({ks[0][1]}) "
          f"p10 {np.percentile([k for k, _ in ks], 10):.4f} median {np.median([k for k, _ in ks]):.4f}; norm-param cos median {np.median(ns):.4f} min {min(ns):.4f}")
ME.set_compute_mode("bf16")
