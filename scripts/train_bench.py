"""Res16UNet34C forward + backward (module path, autograd) on one synthetic scene: voxels/s and algorithmic TFLOP/s.

BASELINE.json metric, first clause: "voxels/sec Res16UNet34C fwd+bwd".  bf16 mode = tcgen05 forward, dgrad and wgrad with fp32
accumulation; activations at module boundaries stay fp32 as in the reference (MinkowskiEngine modules exchange fp32 .F).

usage: python scripts/train_bench.py [scene=s200k|floor2|floor4] [reps] [mode=bf16|fp32] [path=module|fused]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from room_instance_segmentation_mask3d_b200 import lib, mask3d_model, me_api, ops, scenes, weights  # noqa: E402


def backbone_flops(mgr, model, keys):
    """Algorithmic forward FLOPs 2 * P * Cin * Cout over every conv the backbone executes (P = actual pair count)."""
    fl = 0
    P3 = [mgr.neighbor_table(k, k, 3, 1).num_pairs for k in keys]
    N = [mgr.size(k) for k in keys]
    for name, m in model.named_modules():
        if not isinstance(m, (me_api.MinkowskiConvolution, me_api.MinkowskiConvolutionTranspose)) or name == "final":
            continue
        K, cin, cout = (m.kernel.shape if m.kernel.dim() == 3 else (1,) + tuple(m.kernel.shape))
        lvl = m._bench_level
        if K == 27:
            P = P3[lvl]
        elif K == 1:
            P = N[lvl]
        elif K == 8:
            P = N[lvl]  # stride-2 conv / transposed conv: one pair per fine row; lvl = the FINE level
        else:
            P = mgr.neighbor_table(keys[0], keys[0], 5, 1).num_pairs
        fl += 2 * P * cin * cout
    return fl


def tag_levels(model):
    """Level (0 = finest) every conv runs on, from the layer table."""
    model.conv0p1s1._bench_level = 0
    for lvl in range(1, 5):
        getattr(model, f"conv{lvl}p{2 ** (lvl - 1)}s2")._bench_level = lvl - 1  # fine level of the strided conv
        for m in getattr(model, f"block{lvl}").modules():
            if isinstance(m, me_api.MinkowskiConvolution):
                m._bench_level = lvl
    for j, lvl in enumerate(range(4, 0, -1)):
        getattr(model, f"convtr{4 + j}p{2 ** lvl}s2")._bench_level = lvl - 1
        for m in getattr(model, f"block{5 + j}").modules():
            if isinstance(m, me_api.MinkowskiConvolution):
                m._bench_level = lvl - 1
    model.final._bench_level = 0


def run(scene="s200k", reps=5, mode="bf16", quiet=False, path="module"):
    s = {"s200k": lambda: scenes.box_room(), "floor2": lambda: scenes.floor_plan(1000, voxel_m=0.02),
         "floor4": lambda: scenes.floor_plan(1000, voxel_m=0.04)}[scene]()
    c, f = scenes.collate([s])
    c, f = torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda()
    lib.load()
    lib.reserve_device_pool(48.0)
    ops.set_compute_mode(mode)
    model = mask3d_model.Res16UNet34C()
    weights.load_synth(model, seed=0)
    model = model.cuda().train()
    tag_levels(model)
    params = [p for p in model.parameters()]

    def step():
        for p in params:
            p.grad = None
        x = me_api.SparseTensor(features=f, coordinates=c)
        out, fm = model.forward_train(x) if path == "fused" else model(x)
        loss = out.F.float().square().mean() + sum(t.F.float().square().mean() for t in fm[:-1])
        loss.backward()
        return x

    x = step()
    for _ in range(7):  # the caching allocator keeps growing (cudaMalloc) for the first handful of training steps
        step()
    torch.cuda.synchronize()
    mgr = x.coordinate_manager
    keys = [x.coordinate_map_key]
    for _ in range(4):
        keys.append(mgr.stride(keys[-1], 2))
    fl = backbone_flops(mgr, model, keys)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    n0 = c.shape[0]
    res = {"scene": scene, "voxels": n0, "mode": mode, "path": path, "ms_fwd_bwd": ms, "voxels_per_s": n0 / (ms * 1e-3), "fwd_gflop": fl / 1e9,
           "fwd_bwd_tflops_algorithmic": 3 * fl / (ms * 1e-3) / 1e12,
           "includes": "coordinate maps + kernel maps + tile plans + forward + backward (all parameters), one scene per step"}
    if not quiet:
        print(json.dumps(res))
    return res


if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else "s200k", int(sys.argv[2]) if len(sys.argv) > 2 else 5, sys.argv[3] if len(sys.argv) > 3 else "bf16",
        path=sys.argv[4] if len(sys.argv) > 4 else "module")
