"""GPU time of the decoder's query-side CUDA graphs (one replay = post-attention of a level + head of the next)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from room_instance_segmentation_mask3d_b200 import lib, mask3d_model, me_api, ops, scenes, weights  # noqa: E402
c, f = scenes.collate([scenes.floor_plan(1000, voxel_m=0.06)])
c, f = torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda()
model = mask3d_model.Mask3D().eval(); weights.load_synth(model); model = model.cuda()
model(me_api.SparseTensor(features=f, coordinates=c))
qg = next(iter(model._qgraphs.values()))
torch.cuda.synchronize()
for name, fn in (("first (head only)", qg.first), ("layer 0 (post + head)", lambda: qg.layer(0)), ("layer 3", lambda: qg.layer(3))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per replay")
