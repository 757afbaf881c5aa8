#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/cu_diag.txt
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py -x -q > gpurun_out/cu_tests.log 2>&1; tail -3 gpurun_out/cu_tests.log
for c in "96 96 1 bf16" "96 128 1 bf16" "128 128 1 bf16" "128 160 1 f32" "256 128 1 bf16" "128 96 1 bf16 bias" "32 64 1 bf16" "384 256 1 bf16"; do
  for e in 0 64 128 1; do
    MK3D_EPI_STAGE=$e timeout 120 python scripts/conv_diag.py floor2 20 $c 2>&1 | tail -1 >> gpurun_out/cu_diag.txt
  done
done
cat gpurun_out/cu_diag.txt
