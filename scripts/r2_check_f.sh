#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/cf_*.txt
for d in 0 7; do
  MK3D_DEBUG=$d MK3D_TRACE=1 timeout 120 python scripts/conv_diag.py floor2 2 96 96 2>&1 | tail -3 | head -2 | cut -c1-330 >> gpurun_out/cf_diag.txt
done
MK3D_TRACE=1 timeout 120 python scripts/conv_diag.py floor2 2 32 32 2>&1 | tail -3 | head -2 | cut -c1-330 >> gpurun_out/cf_diag.txt
MK3D_TRACE=1 timeout 120 python scripts/conv_diag.py floor2 2 128 128 2>&1 | tail -3 | head -2 | cut -c1-330 >> gpurun_out/cf_diag.txt
cat gpurun_out/cf_diag.txt
