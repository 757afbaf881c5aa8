#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/cj_tests.log 2>&1; tail -5 gpurun_out/cj_tests.log
timeout 900 python bench.py > gpurun_out/cj_bench.json 2> gpurun_out/cj_bench.err; cut -c1-400 gpurun_out/cj_bench.json; tail -3 gpurun_out/cj_bench.err
