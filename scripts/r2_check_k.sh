#!/bin/bash
mkdir -p gpurun_out
for g in 0 6 3; do
  MK3D_DEFER_GROUP=$g timeout 600 python bench.py --steps 8 --no-train --no-cpu-baseline > gpurun_out/ck_bench_$g.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/ck_bench_$g.json'))
print('group $g', 'value', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']/1e6,1), 'other', round(d['e2e']['other_result_form']['value']/1e6,1), 'final_only', round(d['final_prediction_only']['value']/1e6,1))
PY
done
