#!/bin/bash
# wide producer layout (16 gather warps, setmaxnreg): bit-identity against the narrow layout, then the sweep
set -x
mkdir -p gpurun_out
rm -f gpurun_out/cd_*.txt
timeout 300 python scripts/gather_ab.py 8 > gpurun_out/cd_ab.txt 2>&1; grep -c "= 0.000e+00" gpurun_out/cd_ab.txt; grep -v "= 0.000e+00" gpurun_out/cd_ab.txt | head
for c in "96 96" "128 128" "64 64" "256 256" "32 32" "128 96" "96 128"; do
  for m in 0 8; do
    MK3D_GATHER=$m timeout 120 python scripts/conv_diag.py floor2 20 $c 2>&1 | grep -v "mk3d trace" >> gpurun_out/cd_diag.txt
  done
done
for m in 0 8; do MK3D_GATHER=$m timeout 120 python scripts/conv_diag.py floor2 20 96 96 1 >> gpurun_out/cd_diag.txt 2>&1; done
for m in 0 8; do MK3D_GATHER=$m timeout 120 python scripts/conv_diag.py floor2 20 128 32 1 f32 >> gpurun_out/cd_diag.txt 2>&1; done
for m in 0 8; do MK3D_GATHER=$m MK3D_DEBUG=7 timeout 120 python scripts/conv_diag.py floor2 20 96 96 >> gpurun_out/cd_diag.txt 2>&1; done
for m in 0 8; do MK3D_GATHER=$m MK3D_TRACE=1 timeout 120 python scripts/conv_diag.py floor2 2 96 96 2>&1 | tail -2 >> gpurun_out/cd_diag.txt 2>&1; done
cat gpurun_out/cd_diag.txt
