#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/cl_tests.log 2>&1; tail -5 gpurun_out/cl_tests.log
timeout 300 python scripts/profile_step.py 0.02 bf16 8 > /dev/null 2>&1; cp gpurun_out/profile_step_bf16.txt gpurun_out/cl_profile_step_b8.txt; head -14 gpurun_out/cl_profile_step_b8.txt | cut -c1-150
timeout 600 python bench.py --steps 8 --no-train --no-cpu-baseline > gpurun_out/cl_bench.json 2>/dev/null; cut -c1-330 gpurun_out/cl_bench.json
