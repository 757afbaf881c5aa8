#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/cg_*.txt
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q > gpurun_out/cg_tests.log 2>&1; tail -3 gpurun_out/cg_tests.log
for m in 0 8; do
  MK3D_GATHER=$m MK3D_TRACE=1 timeout 120 python scripts/conv_diag.py floor2 2 96 96 2>&1 | tail -3 | head -2 | cut -c1-330 >> gpurun_out/cg_diag.txt
done
for c in "96 96" "128 128" "64 64" "256 256" "32 32" "128 96" "96 128"; do
  for m in 0 8; do
    MK3D_GATHER=$m timeout 120 python scripts/conv_diag.py floor2 20 $c 2>&1 | grep -v "mk3d trace" >> gpurun_out/cg_diag.txt
  done
done
for m in 0 8; do MK3D_GATHER=$m timeout 120 python scripts/conv_diag.py floor2 20 96 96 1 >> gpurun_out/cg_diag.txt 2>&1; done
for m in 0 8; do MK3D_GATHER=$m timeout 120 python scripts/conv_diag.py floor2 20 128 32 1 f32 >> gpurun_out/cg_diag.txt 2>&1; done
for m in 0 8; do MK3D_GATHER=$m MK3D_DEBUG=7 timeout 120 python scripts/conv_diag.py floor2 20 96 96 >> gpurun_out/cg_diag.txt 2>&1; done
cat gpurun_out/cg_diag.txt
