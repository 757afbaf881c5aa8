#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/ch_tests.log 2>&1; tail -3 gpurun_out/ch_tests.log
timeout 900 python bench.py > gpurun_out/ch_bench.json 2> gpurun_out/ch_bench.err; tail -c 1500 gpurun_out/ch_bench.json
timeout 300 python scripts/profile_step.py 0.02 bf16 8 > /dev/null 2>&1; cp gpurun_out/profile_step.txt gpurun_out/ch_profile_step_b8.txt; head -40 gpurun_out/ch_profile_step_b8.txt | cut -c1-150
