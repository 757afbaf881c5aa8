#!/bin/bash
# ablation sweep of k_conv_umma (one process per setting: the switches are read once per process; needs a tuning build:
# MK3D_BUILD_TUNING=1 python -m room_instance_segmentation_mask3d_b200.build --force)
OUT=${OUT:-gpurun_out/diag_r2h.txt}
mkdir -p gpurun_out
: > $OUT
run() { env "$@" python scripts/conv_diag.py floor2 5 $C $C >> $OUT 2>&1; }
for C in 96 128 64 256; do
  run A=0
  run MK3D_TRACE=1
  run MK3D_DEBUG=7
done
grep -v "^\[mk3d trace\]" $OUT | tail -80
grep "^\[mk3d trace\]" $OUT | awk 'NR%3==2' | cut -c1-330
