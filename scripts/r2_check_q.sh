#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/cq_diag.txt
for c in "96 96 1 bf16" "96 128 1 bf16" "128 128 1 bf16" "128 160 1 f32" "128 32 1 f32"; do
  for d in 0 8 1 9; do
    MK3D_DEBUG=$d timeout 120 python scripts/conv_diag.py floor2 20 $c 2>&1 | tail -1 >> gpurun_out/cq_diag.txt
  done
  MK3D_TRACE=1 timeout 120 python scripts/conv_diag.py floor2 2 $c 2>&1 | grep "mk3d trace" | tail -2 | cut -c1-330 >> gpurun_out/cq_diag.txt
done
cat gpurun_out/cq_diag.txt
