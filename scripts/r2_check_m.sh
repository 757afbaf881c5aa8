#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
for c in 1 0; do
  MK3D_KMAP_COMPOSE=$c timeout 600 python bench.py --steps 8 --no-train --no-cpu-baseline > gpurun_out/cm_bench_$c.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/cm_bench_$c.json'))
print('compose $c', 'value', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],1), 'mallocs', d['cuda_mallocs_in_timed_region'], 'e2e', round(d['e2e']['value']/1e6,1), 'other', round(d['e2e']['other_result_form']['value']/1e6,1), 'final_only', round(d['final_prediction_only']['value']/1e6,1))
PY
done
done
