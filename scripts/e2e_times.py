"""Per-step wall times of the e2e step (H2D from pinned memory, forward, D2H of results) + allocator counters."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from room_instance_segmentation_mask3d_b200 import lib, mask3d_model, me_api, ops, scenes, weights  # noqa: E402
dev = torch.device("cuda", 0)
if len(sys.argv) > 2:
    lib.reserve_device_pool(float(sys.argv[2]), dev)
hs = [scenes.collate([scenes.floor_plan(1000 + i, voxel_m=0.02)]) for i in range(3)]
pinned = [(torch.from_numpy(c).pin_memory(), torch.from_numpy(f).pin_memory()) for c, f in hs]
nvox = [c.shape[0] for c, _ in hs]
model = mask3d_model.Mask3D().eval(); weights.load_synth(model); model = model.to(dev)
h_cls = torch.empty((13, 1, 25, 2), dtype=torch.float32).pin_memory()
h_mask = [torch.empty((n, 25), dtype=torch.float32).pin_memory() for n in nvox]
mode = sys.argv[1] if len(sys.argv) > 1 else "async"


state = {"x": None}
lanes = [torch.cuda.Stream(), torch.cuda.Stream()]
for st_ in lanes:
    lib.reserve_device_pool(12.0, dev, stream=st_)
lib.reserve_device_pool(6.0, dev, stream=model.side_stream(dev))
h_cls2 = [torch.empty((13, 1, 25, 2), dtype=torch.float32).pin_memory() for _ in range(2)]
h_mask2 = [[torch.empty((n, 25), dtype=torch.float32).pin_memory() for n in nvox] for _ in range(2)]


def step(i):
    x = state["x"] if state["x"] is not None else model.prepare(*pinned[i % 3])
    ln = i % 2
    with torch.cuda.stream(lanes[ln]):
        o = model(x, prefetch=pinned[(i + 1) % 3])
        h_cls2[ln].copy_(torch.stack(o["pred_class_all_decoders"]), non_blocking=True)
        h_mask2[ln][i % 3].copy_(o["pred_mask_all_decoders"][-1], non_blocking=True)
    state["x"] = o.pop("prefetched")


import gc
if os.environ.get("NOGC"):
    gc.collect(); gc.freeze(); gc.disable()
ts, allocs, mem = [], [], []
torch.cuda.synchronize(); t_prev = time.perf_counter()
for i in range(60):
    step(i)
    if mode == "sync":
        torch.cuda.synchronize()
    t = time.perf_counter(); ts.append((t - t_prev) * 1e3); t_prev = t
    allocs.append(torch.cuda.memory_stats().get("num_device_alloc"))
    mem.append(torch.cuda.memory_allocated() >> 20)
torch.cuda.synchronize()
print(mode, " ".join(f"{t:.1f}" for t in ts))
print("cudaMalloc count per step:", allocs)
print("allocated MiB per step:", mem, "gc counts", gc.get_count())
