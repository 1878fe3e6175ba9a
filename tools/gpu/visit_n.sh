#!/bin/bash
# A/B on one box: in-tree library (A) against tools/gpu/ab/libB.so (B), thread-count variants; parity of A first
TAG=${1:-r02n}
mkdir -p gpurun_out
LIB=flatland-marl_b200/csrc/libflatland_b200.so
cp $LIB /tmp/libA.so
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_deep.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/pytest_$TAG.log 2>&1; tail -5 gpurun_out/pytest_$TAG.log
OUT=gpurun_out/sweep_$TAG.txt; : > $OUT
run() { cfg=$1; shift
  timeout 300 python bench.py --config $cfg --steps 40 --warmup 5 --no-cpu --no-sweep --no-configs --no-policy --episodes 0 --profile-steps 20 "$@" 2>> gpurun_out/sweep_$TAG.err |
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
for rep in 1 2; do
for v in A B; do
  if [ $v = A ]; then cp /tmp/libA.so $LIB; else cp tools/gpu/ab/libB.so $LIB; fi
  for nt in 128 160; do echo "== Test_03 lib=$v nt=$nt (rep $rep)" | tee -a $OUT; FL_OBS_NT=$nt run Test_03 --e2e-steps 3 | tee -a $OUT; [ $rep = 1 ] && FL_OBS_NT=$nt counters Test_03 | tee -a $OUT; done
done; done
cp /tmp/libA.so $LIB
echo "== Test_03 lib=A default" | tee -a $OUT; run Test_03 --e2e-steps 30 | tee -a $OUT
for cfg in Test_02 Test_08 Test_14; do echo "== $cfg lib=A default" | tee -a $OUT; run $cfg --e2e-steps 3 | tee -a $OUT; done
echo "== Test_08 lib=A parts=0" | tee -a $OUT; FL_OBS_PARTS=0 run Test_08 --e2e-steps 3 | tee -a $OUT
cp tools/gpu/ab/libB.so $LIB
echo "== Test_08 lib=B parts=0" | tee -a $OUT; FL_OBS_PARTS=0 run Test_08 --e2e-steps 3 | tee -a $OUT
cp /tmp/libA.so $LIB
