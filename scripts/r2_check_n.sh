#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
  PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True timeout 600 python bench.py --steps 8 --no-train --no-cpu-baseline > gpurun_out/cn_bench.json 2>gpurun_out/cn_bench.err
  python - <<PY
import json
d=json.load(open('gpurun_out/cn_bench.json'))
print('expandable run $i', 'value', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],1), 'mallocs', d['cuda_mallocs_in_timed_region'], 'e2e', round(d['e2e']['value']/1e6,1), 'other', round(d['e2e']['other_result_form']['value']/1e6,1), 'final_only', round(d['final_prediction_only']['value']/1e6,1))
PY
done
tail -3 gpurun_out/cn_bench.err
