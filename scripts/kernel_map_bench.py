import os, sys, torch
sys.path.insert(0, os.getcwd())
from room_instance_segmentation_mask3d_b200 import engine, lib, scenes
lib.load()
c, _ = scenes.collate([scenes.floor_plan(1000, voxel_m=0.02)])
c = torch.from_numpy(c).cuda()
mgr = engine.CoordinateManager(D=3)
k0, _ = mgr.insert_and_map(c)
cm = mgr._maps[k0]
for ks in (3, 5):
    K = ks ** 3
    offs = mgr._offsets_dev((ks,) * 3, (1, 1, 1), k0.tensor_stride)
    a = torch.empty((K, cm.n), dtype=torch.int32, device="cuda")
    for name in ("mk_kernel_map", "mk_kernel_map_same"):
        f = lambda: lib.call(name, lib.p(cm.keys), lib.p(cm.vals), cm.cap, lib.p(cm.coords), cm.n, lib.p(offs), K, lib.p(a), lib.stream())
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        print(f"{ks}^3 {name}: {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
