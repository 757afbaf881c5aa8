#!/bin/bash
# ncu evidence for every kernel family of the path (one gpurun call; the plain run of each command precedes its ncu run)
set -x
mkdir -p gpurun_out
# (1) the dominant kernel, --set full with source: 3^3 conv 96 -> 96 on the 889 k-voxel level-0 map
true && \
ncu --set full --clock-control none --import-source on -k regex:k_conv_umma -s 1 -c 1 -o gpurun_out/r02_conv96_final python scripts/conv_diag.py floor2 2 96 96 > gpurun_out/ncu_c96_final.log 2>&1
# (2) every mk:: kernel of ONE steady-state inference step (SpeedOfLight + memory workload + launch statistics)
python scripts/profile_once.py infer floor2 1 > gpurun_out/plain_infer.log 2>&1 && \
ncu --profile-from-start off --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none \
    -k regex:^k_ --csv --page raw --log-file gpurun_out/r02_families_infer.csv python scripts/profile_once.py infer floor2 1 > gpurun_out/ncu_infer.log 2>&1
# (3) the same for one training step (forward_train + SetCriterion + backward)
python scripts/profile_once.py train floor2 1 > gpurun_out/plain_train.log 2>&1 && \
ncu --profile-from-start off --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none \
    -k regex:^k_ --csv --page raw --log-file gpurun_out/r02_families_train.csv python scripts/profile_once.py train floor2 1 > gpurun_out/ncu_train.log 2>&1
ls -la gpurun_out/r02_families_*.csv gpurun_out/r02_conv96_final.ncu-rep
tail -2 gpurun_out/ncu_infer.log gpurun_out/ncu_train.log
