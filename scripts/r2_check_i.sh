#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/ci_tests.log 2>&1; tail -3 gpurun_out/ci_tests.log
timeout 300 python scripts/profile_step.py 0.02 bf16 8 > /dev/null 2>&1; cp gpurun_out/profile_step_bf16.txt gpurun_out/ci_profile_step_b8.txt; head -16 gpurun_out/ci_profile_step_b8.txt | cut -c1-150
timeout 900 python bench.py > gpurun_out/ci_bench.json 2> gpurun_out/ci_bench.err; cut -c1-400 gpurun_out/ci_bench.json
