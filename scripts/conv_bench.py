"""Sparse-conv microbenchmark (BASELINE.json configs[4]): kernel-map build and tcgen05 gather-GEMM over voxel counts and
channel widths.  Prints one line per case: algorithmic TFLOP/s = 2 * P * Cin * Cout / t (P = actual pair count).

usage: python scripts/conv_bench.py [scene=s200k|floor2|floor4|vox<cm>|blob<side>] [reps] [widths...]
  vox<cm>    the floor-plan generator at <cm> voxels (vox15 = 15 k voxels ... vox1.2 = 2.4 M): the N sweep of SURVEY section 8d config 5
  blob<side> a dense <side>^3 cube (26 neighbours per interior voxel): the dense-occupancy variant"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from room_instance_segmentation_mask3d_b200 import engine, lib, ops, scenes  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "s200k"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
widths = [int(w) for w in sys.argv[3:]] or [32, 64, 96, 128, 256]
if scene.startswith("blob"):
    side = int(scene[4:])
    g = torch.arange(side, dtype=torch.int32)
    xyz = torch.stack(torch.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    c = torch.cat([torch.zeros(len(xyz), 1, dtype=torch.int32), xyz], 1).cuda()
else:
    if scene.startswith("vox"):
        s = scenes.floor_plan(1000, voxel_m=float(scene[3:]) / 100.0)
    else:
        s = {"s200k": lambda: scenes.box_room(), "floor2": lambda: scenes.floor_plan(1000, voxel_m=0.02),
             "floor4": lambda: scenes.floor_plan(1000, voxel_m=0.04)}[scene]()
    c, _ = scenes.collate([s])
    c = torch.from_numpy(c).cuda()
lib.load()


def timeit(fn, n=reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


mgr = engine.CoordinateManager(D=3)
t0 = time.perf_counter()
k0, _ = mgr.insert_and_map(c)
torch.cuda.synchronize()
t_hash = time.perf_counter() - t0
k1 = mgr.stride(k0, 2)
km3 = mgr.neighbor_table(k0, k0, 3, 1)
km2 = mgr.neighbor_table(k0, k1, 2, 1)
tp = mgr.transpose_plan(k1, k0, 2, 1)
n0, n1 = mgr.size(k0), mgr.size(k1)
print(f"scene {scene}: N0={n0} N1={n1} P3={km3.num_pairs} ({km3.num_pairs / n0:.2f}/voxel) P2={km2.num_pairs}; first hash build {t_hash * 1e3:.2f} ms (incl. sync)")


def build_maps():
    m = engine.CoordinateManager(D=3)
    a, _ = m.insert_and_map(c)
    m.neighbor_table(a, a, 3, 1)


print(f"hash build + 3^3 kernel map: {timeit(build_maps, 5):.3f} ms")
sp = km3.sorted_plan()
pop = lambda t: int(sum(bin(int(v) & 0xFFFFFFFF).count('1') for v in t.cpu().numpy()))
print(f"tile efficiency (pairs / (128 * active tile-offsets)): lexicographic {km3.num_pairs / (128 * pop(km3.tile_kmask)):.3f}, "
      f"mask-sorted {km3.num_pairs / (128 * pop(sp.tile_kmask)):.3f}")
for C in widths:
    x = torch.randn(n0, C, device="cuda").bfloat16()
    W = torch.randn(27, C, C, device="cuda") / (27 * C) ** 0.5
    for sort in (False, True):
        ops.SORT_TILES = sort
        t = timeit(lambda: ops.conv_forward(x, W, km3, mode="bf16", out_dtype=torch.bfloat16))
        fl = 2 * km3.num_pairs * C * C
        print(f"conv 3^3 {C:3d}->{C:3d} N={n0} sorted={int(sort)}: {t:.3f} ms  {fl / t / 1e9:8.1f} TFLOP/s")
    ops.SORT_TILES = True
    if os.environ.get("MK3D_BENCH_BWD", "1") != "0":
        dy = torch.randn(n0, C, device="cuda").bfloat16()
        kidx = torch.arange(26, -1, -1, device="cuda", dtype=torch.int32)
        fl = 2 * km3.num_pairs * C * C
        t = timeit(lambda: ops.conv_dgrad_bf16(dy, W, km3, kidx=kidx))
        print(f"dgrad 3^3 {C:3d}->{C:3d} N={n0}: {t:.3f} ms  {fl / t / 1e9:8.1f} TFLOP/s")
        t = timeit(lambda: ops.conv_wgrad_bf16(x, dy, km3, 27, C, C))
        print(f"wgrad 3^3 {C:3d}->{C:3d} N={n0}: {t:.3f} ms  {fl / t / 1e9:8.1f} TFLOP/s")
    W2 = torch.randn(8, C, C, device="cuda") / (8 * C) ** 0.5
    t = timeit(lambda: ops.conv_forward(x, W2, km2, mode="bf16", out_dtype=torch.bfloat16))
    print(f"conv 2^3 s2 {C:3d}->{C:3d} N={n0}->{n1}: {t:.3f} ms  {2 * km2.num_pairs * C * C / t / 1e9:8.1f} TFLOP/s")
    x1 = torch.randn(n1, C, device="cuda").bfloat16()
    t = timeit(lambda: ops.conv_forward(x1, W2, tp, mode="bf16", out_dtype=torch.bfloat16))
    print(f"convT 2^3 s2 {C:3d}->{C:3d} N={n1}->{n0}: {t:.3f} ms  {2 * n0 * C * C / t / 1e9:8.1f} TFLOP/s")
    W1 = torch.randn(1, C, C, device="cuda") / C ** 0.5
    km1 = mgr.neighbor_table(k0, k0, 1, 1)
    t = timeit(lambda: ops.conv_forward(x, W1, km1, mode="bf16", out_dtype=torch.bfloat16))
    print(f"conv 1x1 {C:3d}->{C:3d} N={n0}: {t:.3f} ms  {2 * n0 * C * C / t / 1e9:8.1f} TFLOP/s  ({n0 * C * 4 / t / 1e6:.0f} GB/s in+out)")
