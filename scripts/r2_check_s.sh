#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ct_diag.txt
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py -x -q > gpurun_out/ct_tests.log 2>&1; tail -3 gpurun_out/ct_tests.log
for c in "96 96 1 bf16" "96 128 1 bf16" "128 128 1 bf16" "128 160 1 f32" "256 128 1 bf16" "128 32 1 f32" "128 96 1 bf16 bias" "96 96" "128 128" "256 256" "64 64" "32 32"; do
  timeout 120 python scripts/conv_diag.py floor2 20 $c 2>&1 | tail -1 >> gpurun_out/ct_diag.txt
done
MK3D_TRACE=1 timeout 120 python scripts/conv_diag.py floor2 2 96 128 1 bf16 2>&1 | grep "mk3d trace" | tail -2 | cut -c1-330 >> gpurun_out/ct_diag.txt
cat gpurun_out/ct_diag.txt
