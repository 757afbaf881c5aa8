"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement of the reference's furthest-point-sampling CUDA kernel
(reference: mask3d/third_party/pointnet2/_ext_src/src/sampling_gpu.cu:62-176, wrapper
sampling.cpp:67-88, block size rule include/cuda_utils.h:17-21), including its tie-breaking,
which matters because voxel coordinates are integers and equal distances are common:

  * thread `t` of a `B`-thread block scans points k = t, t+B, ...; keeps the FIRST maximum
    (strict `d2 > best`, lines 111-112), starting from best = -1, besti = 0;
  * points with |p|^2 <= 1e-3 are skipped entirely (line 104): they neither update `temp`
    nor become candidates;
  * the shared-memory tree keeps the left operand on ties (`v2 > v1 ? i2 : i1`, lines 62-68);
  * `temp` starts at 1e10 (sampling.cpp:76-78); the first sample is index 0.

parity pinned: the source of this kernel IS in /root/reference, so the restatement is checked
line by line against it; it cannot be executed here (no GPU, needs ATen build), see DESIGN.md.
"""
from __future__ import annotations

import numpy as np


def fps_block_size(n: int) -> int:
    """opt_n_threads (cuda_utils.h:17-21): min(2^floor(log2 n), 512), at least 1."""
    pow2 = int(np.log(float(n)) / np.log(2.0))
    return max(min(1 << pow2, 512), 1)


def furthest_point_sample(xyz: np.ndarray, npoint: int) -> np.ndarray:
    """xyz float32 [N,3] -> int32 [npoint] indices, identical to the reference kernel."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    n = xyz.shape[0]
    B = fps_block_size(n)
    idxs = np.zeros(npoint, dtype=np.int32)
    temp = np.full(n, 1e10, dtype=np.float32)
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    mag = (x * x) + (y * y) + (z * z)  # fp32, same association as the kernel
    valid = ~(mag <= np.float32(1e-3))
    rows = (n + B - 1) // B
    pad = rows * B - n
    old = 0
    for j in range(1, npoint):
        dx, dy, dz = x - x[old], y - y[old], z - z[old]
        d = (dx * dx + dy * dy) + dz * dz
        d2 = np.minimum(d, temp)
        temp = np.where(valid, d2, temp).astype(np.float32)
        cand = np.where(valid, d2, np.float32(-np.inf)).astype(np.float32)
        if pad:
            cand = np.concatenate([cand, np.full(pad, -np.inf, dtype=np.float32)])
        grid = cand.reshape(rows, B)  # grid[r, t] = point r*B + t, thread t scans column t
        arg = np.argmax(grid, axis=0)  # first maximum in scan order == strict '>' update
        best = grid[arg, np.arange(B)]
        besti = (arg * B + np.arange(B)).astype(np.int32)
        nothing = ~(best > np.float32(-1.0))  # a thread that never beat best=-1 keeps besti=0
        best = np.where(nothing, np.float32(-1.0), best)
        besti = np.where(nothing, 0, besti).astype(np.int32)
        s = B // 2
        while s >= 1:  # shared-memory tree, left operand wins ties
            v1, v2 = best[:s], best[s:2 * s]
            i1, i2 = besti[:s], besti[s:2 * s]
            take2 = v2 > v1
            best = np.where(take2, v2, v1)
            besti = np.where(take2, i2, i1)
            s //= 2
        old = int(besti[0])
        idxs[j] = old
    return idxs
