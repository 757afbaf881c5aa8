#!/bin/bash
# ablation sweep of k_conv_umma (one process per setting: the switches are read once per process)
OUT=gpurun_out/diag_r2a.txt
mkdir -p gpurun_out
: > $OUT
run() { env "$@" python scripts/conv_diag.py floor2 5 $C $C >> $OUT 2>&1; }
for C in 96 64 128; do
  run A=0
  run MK3D_TRACE=1
  for d in 1 2 3 4 5 6 7; do run MK3D_DEBUG=$d; done
  run MK3D_DEBUG=7 MK3D_TRACE=1
  run MK3D_GROUP=1
  run MK3D_STAGES=3
  run MK3D_PRODUCERS=3
  run MK3D_PRODUCERS=2
  run MK3D_CHUNK_MAJOR=0
done
C=256; run A=0; run MK3D_TRACE=1; run MK3D_DEBUG=7; run MK3D_DEBUG=3
C=32; run A=0; run MK3D_TRACE=1; run MK3D_DEBUG=7
cat $OUT | grep -v "^\[mk3d trace\]" | tail -80
