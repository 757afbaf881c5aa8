#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "kernel_map or composed" > gpurun_out/cw_tests.log 2>&1; tail -3 gpurun_out/cw_tests.log
timeout 300 python scripts/profile_step.py 0.02 bf16 8 > /dev/null 2>&1; head -16 gpurun_out/profile_step_bf16.txt | cut -c1-150
