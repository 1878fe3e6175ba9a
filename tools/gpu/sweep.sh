#!/bin/bash
# tuning sweep of the k_observe launch shape (threads per CTA, CTAs per SM planned for, tables in shared memory)
mkdir -p gpurun_out
: > gpurun_out/sweep.txt
for cfg in ${CONFIGS:-Test_03}; do
for nt in ${NTS:-128 256}; do for c in ${CTAS:-8 6 5 4 3}; do for tb in ${TABLES:-63}; do
  FL_OBS_NT=$nt FL_OBS_CTAS=$c FL_OBS_TABLES=$tb timeout 300 python bench.py --config $cfg --steps 60 --warmup 5 --no-cpu --e2e-steps 3 --profile-steps 20 > gpurun_out/sw.json 2> gpurun_out/sw.err
  python - "$cfg" "nt$nt" "$c" "$tb" >> gpurun_out/sweep.txt <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/sw.json'))
    print(sys.argv[1], sys.argv[2],'ctas',sys.argv[3],'tables',sys.argv[4],'value %.1fM'%(d['value']/1e6),'ms/step %.3f'%d['ms_per_step'], {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})
except Exception as e:
    print(sys.argv[1:], 'FAILED', e, open('gpurun_out/sw.err').read()[-300:])
PY
done; done; done; done
cat gpurun_out/sweep.txt
