"""ORACLE BUILD RECIPE — TEST INFRASTRUCTURE ONLY (never imported by the product package).

Stages what of the reference can run in this environment under oracle/_ref/ (git-ignored; it travels to the GPU box with the
snapshot, where /root/reference does not exist).  Needs /root/reference; a no-op without it.

  1. oracle/_ref/mask3d/...   byte-for-byte copies of the reference's OWN Python model code (mask3d/models, utils/measure_runtime.py,
     third_party/pointnet2/{pointnet2_utils,pytorch_utils}.py): tests/boundary_check.py imports it UNMODIFIED on top of the
     drop-in shims of room_instance_segmentation_mask3d_b200/compat (the boundary proof).
  2. oracle/_ref/libfps_ref.so   the reference's furthest-point-sampling CUDA kernel
     (mask3d/third_party/pointnet2/_ext_src/src/sampling_gpu.cu) compiled for sm_100a from the source where it lies, with a stub
     for the two ATen includes of its cuda_utils.h (oracle/ref_stub) and a C entry point (oracle/fps_ref_entry.cu): the REAL
     reference kernel that pins csrc/fps.cu's tie-breaking on the GPU box (tests/test_gpu_boundary.py).

MinkowskiEngine itself is not in /root/reference (un-vendored, pinned 0.5.4) and cannot be staged.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/mask3d"
OUT = os.path.join(HERE, "_ref")
PY_PARTS = ["models", "utils/__init__.py", "utils/measure_runtime.py", "third_party/pointnet2/pointnet2_utils.py",
            "third_party/pointnet2/pytorch_utils.py"]


def stage_python() -> str:
    dst_root = os.path.join(OUT, "mask3d")
    if os.path.isdir(dst_root):
        for d, _, _f in os.walk(dst_root):
            os.chmod(d, 0o755)
        shutil.rmtree(dst_root)
    for part in PY_PARTS:
        src, dst = os.path.join(REF, part), os.path.join(dst_root, part)
        if os.path.isdir(src):
            shutil.copytree(src, dst, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.exists(src):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
    for d, _, files in os.walk(dst_root):  # the reference tree is read-only; the staged copy must stay replaceable
        os.chmod(d, 0o755)
        for f in files:
            os.chmod(os.path.join(d, f), 0o644)
    return dst_root


def build_fps() -> str:
    lib = os.path.join(OUT, "libfps_ref.so")
    src = os.path.join(REF, "third_party/pointnet2/_ext_src/src/sampling_gpu.cu")
    inc = os.path.join(REF, "third_party/pointnet2/_ext_src/include")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
           "-I", os.path.join(HERE, "ref_stub"), "-I", inc, src, os.path.join(HERE, "fps_ref_entry.cu"), "-o", lib, "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the reference FPS kernel failed:\n" + r.stdout)
    return lib


def main() -> int:
    if not os.path.isdir(REF):
        print("oracle/make_ref.py: /root/reference not present, nothing staged")
        return 0
    os.makedirs(OUT, exist_ok=True)
    print("staged", stage_python())
    print("built", build_fps())
    return 0


if __name__ == "__main__":
    sys.exit(main())
