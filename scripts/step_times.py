"""Per-step wall time of the bench step over the first N steps (allocator warm-up study)."""
import os, sys, time
if len(sys.argv) > 1 and sys.argv[1] == "expandable":
    os.environ["PYTORCH_CUDA_ALLOC_CONF"] = "expandable_segments:True"
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from room_instance_segmentation_mask3d_b200 import lib, mask3d_model, me_api, ops, scenes, weights  # noqa: E402
hs = [scenes.collate([scenes.floor_plan(1000 + i, voxel_m=0.02)]) for i in range(3)]
res = [(torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda()) for c, f in hs]
model = mask3d_model.Mask3D().eval(); weights.load_synth(model); model = model.cuda()
ts = []
for i in range(24):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c, f = res[i % 3]
    model(me_api.SparseTensor(features=f, coordinates=c))
    torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print(os.environ.get("PYTORCH_CUDA_ALLOC_CONF", "default"), " ".join(f"{t:.1f}" for t in ts))
st = torch.cuda.memory_stats()
print("num_alloc_retries", st.get("num_alloc_retries"), "segments", st.get("segment.all.current"), "reserved GB", st.get("reserved_bytes.all.current") / 1e9,
      "cudaMalloc count", st.get("num_device_alloc"), "free", st.get("num_device_free"))
