#!/bin/bash
# parity tests, phase times of the headline config, and a short bench of every BASELINE.json config
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/phase_times.py Test_03 > gpurun_out/phase.txt 2>&1; cat gpurun_out/phase.txt
: > gpurun_out/configs.txt
for cfg in Test_03 Test_02 Test_08 Test_14 Test_00; do
  timeout 600 python bench.py --config $cfg --steps 60 --warmup 5 --no-cpu --e2e-steps 10 --profile-steps 20 > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  python - $cfg >> gpurun_out/configs.txt <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/bench_%s.json'%sys.argv[1]))
    print(sys.argv[1],'value %.1fM'%(d['value']/1e6),'ms/step %.3f'%d['ms_per_step'],'e2e %.1fM'%(d['e2e']['value']/1e6),{k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()}, 'envs',d['config']['envs_per_gpu'], 'agents',d['config']['agents_per_env'])
except Exception as e:
    print(sys.argv[1],'FAILED',e,open('gpurun_out/bench_%s.err'%sys.argv[1]).read()[-400:])
PY
done
cat gpurun_out/configs.txt
