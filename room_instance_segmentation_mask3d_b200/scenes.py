"""Synthetic Structured3D-shaped room scenes (integer voxel coordinates).

The reference trains/evaluates on voxelised Structured3D rooms whose features are a
single channel of ones (reference: mask3d/datasets/semseg_structured3d.py:174) and whose
coordinates are integer voxel indices, deduplicated and lexicographically sorted per scene
by ``np.unique(axis=0)`` (reference: mask3d/datasets/utils.py:106).  There is no dataset on
the bench box, so every test / bench input comes from the generators below.  They rasterise
surfaces (floor, ceiling, walls, furniture boxes) directly on the voxel grid, then apply the
same ``np.unique`` so row order matches what the reference's collate would hand to
``ME.SparseTensor``.

numpy only: this file is shared by the product, the oracle tests and bench.py.
"""
from __future__ import annotations

import numpy as np


def _plane(ax: int, val: int, r0, r1) -> np.ndarray:
    """All voxels of an axis-aligned rectangle: coordinate `ax` fixed to `val`,
    the two other axes spanning [r0[i], r1[i]) in axis order."""
    others = [a for a in range(3) if a != ax]
    u = np.arange(r0[0], r1[0], dtype=np.int32)
    v = np.arange(r0[1], r1[1], dtype=np.int32)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    out = np.empty((uu.size, 3), dtype=np.int32)
    out[:, ax] = val
    out[:, others[0]] = uu.ravel()
    out[:, others[1]] = vv.ravel()
    return out


def _box_surface(lo, hi) -> np.ndarray:
    """Surface voxels of the axis-aligned box [lo, hi) (6 faces)."""
    lo = np.asarray(lo, dtype=np.int64)
    hi = np.asarray(hi, dtype=np.int64)
    faces = []
    for ax in range(3):
        others = [a for a in range(3) if a != ax]
        r0 = (lo[others[0]], lo[others[1]])
        r1 = (hi[others[0]], hi[others[1]])
        faces.append(_plane(ax, int(lo[ax]), r0, r1))
        faces.append(_plane(ax, int(hi[ax] - 1), r0, r1))
    return np.concatenate(faces, axis=0)


def _finish(pts: np.ndarray) -> np.ndarray:
    """np.unique(axis=0) == the reference's per-scene dedup + lexicographic order."""
    return np.unique(pts.astype(np.int32), axis=0)


def box_room(size_m=(4.0, 5.0, 2.8), voxel_m: float = 0.02) -> np.ndarray:
    """Empty box room, surfaces only.  Defaults give the S200k scene of SURVEY.md section 8
    (N0 = 223 648 voxels at 2 cm)."""
    dims = [int(round(s / voxel_m)) for s in size_m]
    return _finish(_box_surface((0, 0, 0), dims))


def floor_plan(
    seed: int = 1000,
    footprint_m=(10.0, 8.0),
    height_m: float = 2.8,
    voxel_m: float = 0.02,
    n_furniture=(5, 10),
    z0_vox: int = 0,
) -> np.ndarray:
    """Single-floor, 6-room apartment: a 3 x 2 grid of rooms separated by axis-aligned
    interior walls with 0.9 m door gaps, floor + ceiling + outer walls, and 5-10 furniture
    boxes standing on the floor (Structured3D averages 6.2 rooms
    on 80 m^2, reference: datasets_preprocess/structured3d_analyze/stru3d_analyze_20241019.txt:10,16).
    """
    rng = np.random.default_rng(seed)
    X = int(round(footprint_m[0] / voxel_m))
    Y = int(round(footprint_m[1] / voxel_m))
    Z = int(round(height_m / voxel_m))
    door_w = max(2, int(round(0.9 / voxel_m)))
    door_h = max(2, int(round(2.0 / voxel_m)))
    parts = [_box_surface((0, 0, 0), (X, Y, Z))]

    # interior walls: two planes x = const, one plane y = const, each jittered a little
    xs = [int(X * f + rng.integers(-X // 40, X // 40 + 1)) for f in (1 / 3, 2 / 3)]
    ys = [int(Y * 0.5 + rng.integers(-Y // 40, Y // 40 + 1))]
    x_edges = [0] + xs + [X]
    y_edges = [0] + ys + [Y]
    for xw in xs:
        wall = _plane(0, xw, (1, 1), (Y - 1, Z - 1))
        keep = np.ones(len(wall), dtype=bool)
        # one door per room this wall borders (one per y-segment)
        for j in range(len(y_edges) - 1):
            seg0, seg1 = y_edges[j], y_edges[j + 1]
            if seg1 - seg0 > door_w + 4:
                d0 = int(rng.integers(seg0 + 2, seg1 - door_w - 1))
                keep &= ~((wall[:, 1] >= d0) & (wall[:, 1] < d0 + door_w) & (wall[:, 2] < door_h))
        parts.append(wall[keep])
    for yw in ys:
        wall = _plane(1, yw, (1, 1), (X - 1, Z - 1))
        keep = np.ones(len(wall), dtype=bool)
        for i in range(len(x_edges) - 1):
            seg0, seg1 = x_edges[i], x_edges[i + 1]
            if seg1 - seg0 > door_w + 4:
                d0 = int(rng.integers(seg0 + 2, seg1 - door_w - 1))
                keep &= ~((wall[:, 0] >= d0) & (wall[:, 0] < d0 + door_w) & (wall[:, 2] < door_h))
        parts.append(wall[keep])

    # furniture: boxes on the floor
    nf = int(rng.integers(n_furniture[0], n_furniture[1] + 1))
    for _ in range(nf):
        sx = max(2, int(round(rng.uniform(0.4, 2.0) / voxel_m)))
        sy = max(2, int(round(rng.uniform(0.4, 1.2) / voxel_m)))
        sz = max(2, int(round(rng.uniform(0.4, 1.8) / voxel_m)))
        if X - sx - 2 <= 2 or Y - sy - 2 <= 2:
            continue
        x0 = int(rng.integers(2, X - sx - 2))
        y0 = int(rng.integers(2, Y - sy - 2))
        parts.append(_box_surface((x0, y0, 1), (x0 + sx, y0 + sy, 1 + sz)))

    pts = np.concatenate(parts, axis=0)
    pts[:, 2] += z0_vox
    return _finish(pts)


def building(
    seed: int = 2000,
    floors: int = 4,
    footprint_m=(20.0, 15.0),
    height_m: float = 2.8,
    voxel_m: float = 0.02,
) -> np.ndarray:
    """Multi-floor building (BASELINE.json configs[3]): `floors` stacked floor plans."""
    Z = int(round(height_m / voxel_m))
    parts = [
        floor_plan(seed + f, footprint_m, height_m, voxel_m, n_furniture=(10, 20), z0_vox=f * Z)
        for f in range(floors)
    ]
    return _finish(np.concatenate(parts, axis=0))


def tiny_scene(seed: int = 0, extent=(24, 20, 12), n_boxes: int = 2) -> np.ndarray:
    """Few-thousand-voxel scene for parity tests: a small box room plus boxes, shifted so that
    negative coordinates occur (training augmentation produces them)."""
    rng = np.random.default_rng(seed)
    X, Y, Z = extent
    parts = [_box_surface((0, 0, 0), (X, Y, Z))]
    for _ in range(n_boxes):
        s = rng.integers(2, 6, size=3)
        o = np.array([rng.integers(1, X - s[0] - 1), rng.integers(1, Y - s[1] - 1), 1])
        parts.append(_box_surface(o, o + s))
    pts = np.concatenate(parts, axis=0)
    pts = pts - np.array([X // 3, Y // 2, 2], dtype=np.int32)
    return _finish(pts)


def collate(scenes) -> tuple[np.ndarray, np.ndarray]:
    """Batch-first int32 coordinates [sum N, 4] and ones features [sum N, 1], in list order —
    what ME.utils.sparse_collate returns for the reference's collate (datasets/utils.py:126)."""
    coords = np.concatenate(
        [np.concatenate([np.full((len(s), 1), b, dtype=np.int32), s.astype(np.int32)], axis=1) for b, s in enumerate(scenes)],
        axis=0,
    )
    feats = np.ones((len(coords), 1), dtype=np.float32)
    return coords, feats
