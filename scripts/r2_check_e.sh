#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/ce_*.txt
for d in 0 1 2 3 4 7; do
  MK3D_DEBUG=$d MK3D_TRACE=1 timeout 120 python scripts/conv_diag.py floor2 2 96 96 2>&1 | tail -2 | head -1 | cut -c1-330 >> gpurun_out/ce_diag.txt
  MK3D_DEBUG=$d timeout 120 python scripts/conv_diag.py floor2 20 96 96 2>&1 | tail -1 >> gpurun_out/ce_diag.txt
done
for d in 0 7; do
  for s in 2 3; do MK3D_DEBUG=$d MK3D_STAGES=$s timeout 120 python scripts/conv_diag.py floor2 20 96 96 2>&1 | tail -1 >> gpurun_out/ce_diag.txt; done
  MK3D_DEBUG=$d MK3D_GROUP=1 timeout 120 python scripts/conv_diag.py floor2 20 96 96 2>&1 | tail -1 >> gpurun_out/ce_diag.txt
  MK3D_DEBUG=$d MK3D_SPLIT=4 timeout 120 python scripts/conv_diag.py floor2 20 96 96 2>&1 | tail -1 >> gpurun_out/ce_diag.txt
done
cat gpurun_out/ce_diag.txt
