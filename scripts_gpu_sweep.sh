#!/bin/bash
# parity first, then a tuning sweep of the k_observe launch shape (threads per CTA, CTAs per SM)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
if ! grep -q "pytest rc=0" gpurun_out/pytest_gpu.log; then
  timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/sanitizer.log 2>&1; tail -40 gpurun_out/sanitizer.log
  exit 1
fi
: > gpurun_out/sweep.txt
for cfg in ${CONFIGS:-Test_03}; do
for g in ${GS:-8}; do for nt in ${NTS:-64 128 256}; do for c in ${CTAS:-2 3 4 5 6}; do
  FL_OBS_NT=$nt FL_OBS_CTAS=$c timeout 300 python bench.py --config $cfg --steps 60 --warmup 5 --no-cpu --e2e-steps 3 --profile-steps 20 > gpurun_out/sw.json 2> gpurun_out/sw.err
  python - "$cfg" "g$g-nt$nt" "$c" >> gpurun_out/sweep.txt <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/sw.json'))
    print(sys.argv[1], 'nt',sys.argv[2],'ctas',sys.argv[3],'value %.1fM'%(d['value']/1e6),'ms/step %.3f'%d['ms_per_step'], {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})
except Exception as e:
    print(sys.argv[1:], 'FAILED', e, open('gpurun_out/sw.err').read()[-300:])
PY
done; done; done; done
cat gpurun_out/sweep.txt
