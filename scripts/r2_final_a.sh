#!/bin/bash
# round-2 final single-GPU evidence: smoke, default bench line, reference arm, ncu launch list of the bench command, ncu --set full of the top kernel
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fa_smoke.log 2>&1; tail -2 gpurun_out/fa_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/fa_bench.err; cut -c1-300 gpurun_out/r02_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/fa_ref.err; cut -c1-600 gpurun_out/r02_bench_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-train --no-cpu-baseline > gpurun_out/fa_ncu_bench.log 2>&1; tail -2 gpurun_out/fa_ncu_bench.log | cut -c1-200
python scripts/conv_diag.py floor2 2 96 96 > gpurun_out/fa_plain_c96.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_umma -s 1 -c 1 -o gpurun_out/r02_conv96_final2 python scripts/conv_diag.py floor2 2 96 96 > gpurun_out/fa_ncu_c96.log 2>&1; tail -2 gpurun_out/fa_ncu_c96.log
ls -la gpurun_out/r02_launches_bench.csv gpurun_out/r02_conv96_final2.ncu-rep
