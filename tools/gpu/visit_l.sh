#!/bin/bash
# experiment switches of the tree phase: parity, then per switch the kernel time (bench) and the instruction counters (ncu, 5 metrics)
TAG=${1:-r02l}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "plan or group" > gpurun_out/pytest_$TAG.log 2>&1; tail -5 gpurun_out/pytest_$TAG.log
OUT=gpurun_out/sweep_$TAG.txt; : > $OUT
run() { cfg=$1; shift
  timeout 300 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu --no-sweep --no-configs --no-policy --episodes 0 --profile-steps 10 "$@" 2>> gpurun_out/sweep_$TAG.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1f M  ms/step %.4f  kernels %s  e2e %.1f M  plan %s' % (d['value']/1e6, d['ms_per_step'], {k: round(v['ms_per_launch']*v['launches_per_step'],4) for k,v in d['kernels'].items()}, d['e2e']['value']/1e6, d['config']['obs_plan']))"; }
counters() { cfg=$1
  timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'k_observe' -s 6 -c ${CNT:-1} --csv python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu --e2e-steps 1 --profile-steps 1 --episodes 0 --no-sweep --no-configs --no-policy 2>/dev/null |
    python -c "
import sys,csv
rows=[r for r in csv.reader(sys.stdin) if len(r)>10 and r[0].isdigit()]
out={}
for r in rows: out.setdefault(r[4].split('(')[0][-30:],{})[r[-3]]=r[-1]
for k,v in out.items(): print('   ncu', k, ' '.join('%s=%s' % (a.split('.')[0].replace('smsp__','').replace('sm__',''), b) for a,b in v.items()))"; }
for x in ${EXPS:-0 1 2 3 4 6}; do echo "== Test_03 exp=$x" | tee -a $OUT; FL_OBS_EXP=$x run Test_03 --e2e-steps 3 | tee -a $OUT; FL_OBS_EXP=$x counters Test_03 | tee -a $OUT; done
for x in ${EXPS2:-0 2}; do echo "== Test_14 exp=$x" | tee -a $OUT; FL_OBS_EXP=$x run Test_14 --e2e-steps 3 | tee -a $OUT; done
for x in ${FW:-0 2 3}; do echo "== Test_02 flatwalk=$x" | tee -a $OUT; FL_OBS_FLATWALK=$x run Test_02 --e2e-steps 3 | tee -a $OUT; done
