"""Generate tests/golden/fps_ref_kernel.npz with the REFERENCE's own furthest-point-sampling CUDA kernel
(mask3d/third_party/pointnet2/_ext_src/src/sampling_gpu.cu:72-176), compiled unmodified by oracle/make_ref.py into
oracle/_ref/libfps_ref.so.  Needs a GPU (run on the B200 box; writes gpurun_out/fps_ref_kernel.npz, which is then committed as
tests/golden/fps_ref_kernel.npz):

    python tests/golden/make_golden_fps.py

Cases: integer voxel coordinates (ties are common: the tie-break is part of parity), a cloud with the origin point (the kernel
skips points with |p|^2 <= 1e-3), non-power-of-two and tiny sizes (block size = min(2^floor(log2 n), 512))."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from room_instance_segmentation_mask3d_b200 import scenes  # noqa: E402


def cases():
    rng = np.random.default_rng(7)
    out = {
        "tiny0": scenes.tiny_scene(0).astype(np.float32),
        "tiny1": scenes.tiny_scene(1, (20, 16, 10)).astype(np.float32),
        "box": scenes.box_room((1.0, 1.2, 0.6), 0.02).astype(np.float32),
        "origin": np.concatenate([np.zeros((1, 3), np.float32), scenes.tiny_scene(5, (9, 7, 5), 0).astype(np.float32)]),
        "float1000": rng.normal(size=(1000, 3)).astype(np.float32),
        "float777": (rng.random((777, 3)) * 10).astype(np.float32),
        "grid30": np.stack(np.meshgrid(np.arange(30), np.arange(30), np.arange(2), indexing="ij"), -1).reshape(-1, 3).astype(np.float32),
        "s40": rng.integers(0, 6, size=(40, 3)).astype(np.float32),
    }
    return out


def run_reference_kernel(lib, xyz: np.ndarray, m: int) -> np.ndarray:
    """what sampling.cpp:67-88 does around the kernel: temp = 1e10, idx = zeros int32"""
    n = xyz.shape[0]
    x = torch.from_numpy(np.ascontiguousarray(xyz)).cuda().contiguous()
    temp = torch.full((1, n), 1e10, dtype=torch.float32, device="cuda")
    idx = torch.zeros((1, m), dtype=torch.int32, device="cuda")
    rc = lib.fps_ref(1, n, m, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(temp.data_ptr()), ctypes.c_void_p(idx.data_ptr()),
                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert rc == 0, f"reference FPS kernel failed: cuda error {rc}"
    return idx[0].cpu().numpy()


def main():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libfps_ref.so"))
    lib.fps_ref.restype = ctypes.c_int
    fix = {}
    for name, xyz in cases().items():
        for m in (25, 100):
            if xyz.shape[0] <= m:
                continue
            fix[f"{name}_xyz"] = xyz
            fix[f"{name}_idx{m}"] = run_reference_kernel(lib, xyz, m)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "fps_ref_kernel.npz")
    np.savez_compressed(path, **fix)
    print("wrote", path, sorted(k for k in fix if "idx" in k))


if __name__ == "__main__":
    main()
