"""A/B of the deferred, stacked mask-logit products in the inference forward (batch of 8 floor scenes): per-step time and allocator activity."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from room_instance_segmentation_mask3d_b200 import lib, mask3d_model, me_api, ops, scenes, weights  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ops.set_compute_mode("bf16")
c, f = scenes.collate([scenes.floor_plan(1000 + i, voxel_m=0.02) for i in range(nb)])
c, f = torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda()
model = mask3d_model.Mask3D().eval()
weights.load_synth(model)
model = model.cuda()
for group in (0, 6, 4, 3, 2):
    os.environ["MK3D_DEFER_GROUP"] = str(group)
    for i in range(7):
        torch.cuda.synchronize()
        st0 = torch.cuda.memory_stats()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        o = model(me_api.SparseTensor(features=f, coordinates=c))
        e1.record()
        t_cpu = time.perf_counter() - t0
        torch.cuda.synchronize()
        st1 = torch.cuda.memory_stats()
        del o
        if i >= 3:
            print(f"group {group} step {i}: gpu {e0.elapsed_time(e1):.1f} ms, enqueue {t_cpu * 1e3:.1f} ms, cudaMalloc calls {st1['num_device_alloc'] - st0['num_device_alloc']}, "
                  f"cudaFree {st1['num_device_free'] - st0['num_device_free']}, retries {st1['num_alloc_retries'] - st0['num_alloc_retries']}, "
                  f"reserved {st1['reserved_bytes.all.current'] / 2**30:.1f} GiB, peak alloc {st1['allocated_bytes.all.peak'] / 2**30:.1f} GiB", flush=True)
