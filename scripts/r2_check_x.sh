#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/cx_diag.txt
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py -x -q > gpurun_out/cx_tests.log 2>&1; tail -3 gpurun_out/cx_tests.log
for c in "96 96" "128 128" "64 64" "128 96" "256 256" "96 96 1 bf16" "128 96 1 bf16 bias" "96 96 3 f32"; do
  for e in 0 1; do
    MK3D_VEC32=$e timeout 120 python scripts/conv_diag.py floor2 20 $c 2>&1 | tail -1 >> gpurun_out/cx_diag.txt
  done
done
cat gpurun_out/cx_diag.txt
