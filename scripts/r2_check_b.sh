#!/bin/bash
# hybrid LDGSTS + TMA gather: bit-identity against the LDGSTS-only producer, then the sweep over the TMA share
set -x
mkdir -p gpurun_out
rm -f gpurun_out/cb_*.txt
for m in 3 4; do timeout 300 python scripts/gather_ab.py $m 2>&1 | awk '{print "mode '$m' " $0}' | sort | uniq -c -f8 | head -40 >> gpurun_out/cb_ab.txt; done
cat gpurun_out/cb_ab.txt | cut -c1-150
for c in "96 96" "128 128" "64 64" "256 256" "32 32" "128 96" "96 128"; do
  for m in 0 3 4 5; do
    MK3D_GATHER=$m timeout 120 python scripts/conv_diag.py floor2 20 $c 2>&1 | grep -v "mk3d trace" >> gpurun_out/cb_diag.txt
  done
done
for m in 0 3 4; do MK3D_GATHER=$m timeout 120 python scripts/conv_diag.py floor2 20 96 96 1 >> gpurun_out/cb_diag.txt 2>&1; done
cat gpurun_out/cb_diag.txt
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "group_rows or mask_compact or wgrad" > gpurun_out/cb_tests.log 2>&1; tail -3 gpurun_out/cb_tests.log
