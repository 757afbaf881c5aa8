#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/cc_*.txt
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -x -q > gpurun_out/cc_tests.log 2>&1; tail -3 gpurun_out/cc_tests.log
for c in "96 96" "32 32" "64 64" "128 96" "96 128" "32 64" "128 128"; do
  timeout 120 python scripts/conv_diag.py floor2 20 $c 2>&1 | grep -v "mk3d trace" >> gpurun_out/cc_diag.txt
done
timeout 120 python scripts/conv_diag.py floor2 20 96 96 1 >> gpurun_out/cc_diag.txt 2>&1
timeout 120 python scripts/conv_diag.py floor2 20 32 32 1 >> gpurun_out/cc_diag.txt 2>&1
cat gpurun_out/cc_diag.txt
