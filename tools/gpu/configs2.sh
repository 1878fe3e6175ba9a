#!/bin/bash
# a short bench of every BASELINE.json config including the policy leg, and an e2e chunk sweep on the headline config
mkdir -p gpurun_out
: > gpurun_out/configs_v14.txt
for cfg in Test_03 Test_02 Test_08 Test_14 Test_00; do
  timeout 600 python bench.py --config $cfg --steps 60 --warmup 5 --no-cpu --e2e-steps 10 --profile-steps 20 --policy-steps 20 > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  python - $cfg >> gpurun_out/configs_v14.txt <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/bench_%s.json'%sys.argv[1]))
    p=d['policy']
    print(sys.argv[1],'envs',d['config']['envs_per_gpu'],'agents',d['config']['agents_per_env'],'| step+obs value %.1fM'%(d['value']/1e6),'ms/step %.3f'%d['ms_per_step'],'e2e %.1fM'%(d['e2e']['value']/1e6),
          '| policy forward %.3f ms = %.1fM agent-steps/s, %.0f TFLOP/s (%.1f%% of peak), rollout %.1fM agent-steps/s' % (p['forward_ms'], p['agent_steps_per_s']/1e6, p['roofline']['achieved'], 100*p['roofline']['frac'], p['rollout']['value']/1e6))
except Exception as e:
    print(sys.argv[1],'FAILED',e,open('gpurun_out/bench_%s.err'%sys.argv[1]).read()[-400:])
PY
done
for ch in 2 4 8 16; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-policy --e2e-steps 40 --e2e-chunks $ch --profile-steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('e2e chunks $ch: %.2fM agent-steps/s' % (d['e2e']['value']/1e6))" >> gpurun_out/configs_v14.txt
done
cat gpurun_out/configs_v14.txt
