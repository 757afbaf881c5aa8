#!/bin/bash
mkdir -p gpurun_out
for l in 2 2 2 1 1; do
  MK3D_BENCH_STEP_DEBUG=1 timeout 600 python bench.py --steps 16 --lanes $l --no-train --no-cpu-baseline 2>/dev/null > gpurun_out/cp.out
  grep "step debug" gpurun_out/cp.out | cut -c1-300
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/cp.out') if l.startswith('{')][-1])
print('lanes $l', 'value', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']/1e6,1), round(d['e2e']['ms_per_step'],1), 'final_only', round(d['final_prediction_only']['ms_per_step'],1))
PY
done
