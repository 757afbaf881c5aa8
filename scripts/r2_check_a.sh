#!/bin/bash
# conv / wgrad kernel changes: parity tests, then the forward sweep with both gather mechanisms and the backward microbenchmark
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q > gpurun_out/ca_tests.log 2>&1; tail -3 gpurun_out/ca_tests.log
timeout 300 python scripts/gather_ab.py > gpurun_out/ca_ab.txt 2>&1; cat gpurun_out/ca_ab.txt
for c in "96 96" "128 128" "64 64" "256 256" "32 32"; do
  timeout 120 python scripts/conv_diag.py floor2 20 $c 2>&1 | grep -v "mk3d trace" >> gpurun_out/ca_diag.txt
  MK3D_GATHER=2 timeout 120 python scripts/conv_diag.py floor2 20 $c 2>&1 | grep -v "mk3d trace" >> gpurun_out/ca_diag.txt
done
timeout 120 python scripts/conv_diag.py floor2 20 96 96 1 >> gpurun_out/ca_diag.txt 2>&1
MK3D_GATHER=2 timeout 120 python scripts/conv_diag.py floor2 20 96 96 1 >> gpurun_out/ca_diag.txt 2>&1
cat gpurun_out/ca_diag.txt
MK3D_BENCH_BWD=1 timeout 300 python scripts/conv_bench.py floor2 5 96 128 > gpurun_out/ca_bwd.txt 2>&1; cat gpurun_out/ca_bwd.txt
