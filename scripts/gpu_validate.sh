#!/bin/bash
# One gpurun call that validates a change on a B200: the whole GPU suite, then the default bench line in short.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_validate.sh'
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/validate_tests.log 2>&1; tail -3 gpurun_out/validate_tests.log
timeout 900 python bench.py --steps 16 --no-cpu-baseline > gpurun_out/validate_bench.json 2>/dev/null
python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/validate_bench.json') if l.startswith('{')][-1])
print('value', round(d['value'] / 1e6, 1), 'M voxels/s,', round(d['ms_per_step'], 1), 'ms/step; e2e', round(d['e2e']['value'] / 1e6, 1), '(', d['e2e']['other_result_form']['results'],
      round(d['e2e']['other_result_form']['value'] / 1e6, 1), '); final prediction only', round(d['final_prediction_only']['value'] / 1e6, 1))
r = d['roofline']
print('spatial-conv class frac of tensor peak', round(r['frac'], 3), '; dense class frac of HBM peak', round(r['classes']['K=1_hbm_bound']['frac_of_hbm_peak'], 3))
print('by shape ms/step', r['by_shape_ms_per_step'])
print('train fwd+bwd ms', round(d['train_fwd_bwd']['ms_fwd_bwd'], 1), '; train step ms', round(d['train_step']['ms_step'], 1))
PY
