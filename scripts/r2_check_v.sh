#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/cv_tests.log 2>&1; tail -3 gpurun_out/cv_tests.log
timeout 900 python bench.py --steps 16 --no-cpu-baseline > gpurun_out/cv_bench.json 2>/dev/null
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/cv_bench.json') if l.startswith('{')][-1])
print('value', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']/1e6,1), 'other', round(d['e2e']['other_result_form']['value']/1e6,1), 'final_only', round(d['final_prediction_only']['value']/1e6,1))
r=d['roofline']; print(r['frac']); print(r['classes']['K=1_hbm_bound']); print(r['by_shape_ms_per_step']); print(d['train_fwd_bwd']['ms_fwd_bwd'], d['train_step']['ms_step'])
PY
