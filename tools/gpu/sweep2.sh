#!/bin/bash
# k_observe launch-shape sweep at the headline config (tuning)
mkdir -p gpurun_out; : > gpurun_out/sweep_v16.txt
for nt in 96 128 160 192 256; do for ctas in "" 5 6 7 8 10; do
  v="FL_OBS_NT=$nt"; [ -n "$ctas" ] && v="$v FL_OBS_CTAS=$ctas"
  env $v timeout 120 python bench.py --steps 60 --warmup 10 --no-cpu --no-policy --e2e-steps 3 --profile-steps 20 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$v','value %.1fM'%(d['value']/1e6), {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})
except Exception as e: print('$v failed', e)" >> gpurun_out/sweep_v16.txt
done; done
cat gpurun_out/sweep_v16.txt
