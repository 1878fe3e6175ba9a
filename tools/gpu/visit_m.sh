#!/bin/bash
# parity of all plans, then: default vs 160 / 192 threads per environment (ncu counters beside the bench), hot-lines profile
TAG=${1:-r02m}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_deep.py -m gpu -q -x > gpurun_out/pytest_$TAG.log 2>&1; tail -5 gpurun_out/pytest_$TAG.log
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
echo "== Test_03 default" | tee -a $OUT; run Test_03 --e2e-steps 3 | tee -a $OUT; counters Test_03 | tee -a $OUT
for nt in 160 192; do echo "== Test_03 nt=$nt" | tee -a $OUT; FL_OBS_NT=$nt run Test_03 --e2e-steps 3 | tee -a $OUT; FL_OBS_NT=$nt counters Test_03 | tee -a $OUT; done
for ss in 12 20; do echo "== Test_03 sortsmall=$ss" | tee -a $OUT; FL_OBS_SORTSMALL=$ss run Test_03 --e2e-steps 3 | tee -a $OUT; done
for cfg in Test_02 Test_08 Test_14; do echo "== $cfg default" | tee -a $OUT; run $cfg --e2e-steps 3 | tee -a $OUT; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_observe' -s 6 -c 1 -o gpurun_out/ncu_Test_03_$TAG -f \
    python bench.py --config Test_03 --steps 3 --warmup 3 --no-cpu --e2e-steps 1 --profile-steps 1 --episodes 0 --no-sweep --no-configs --no-policy > gpurun_out/ncu_full_Test_03_$TAG.log 2>&1
ncu -i gpurun_out/ncu_Test_03_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_Test_03_${TAG}_raw.csv 2>/dev/null
NCU_KRE=k_observe python profiles/hot_lines.py gpurun_out/ncu_Test_03_$TAG.ncu-rep 'k_observeILi128ELi7ELi0' 70 > gpurun_out/hot_lines_Test_03_$TAG.txt 2>&1
python profiles/summarize.py gpurun_out/ncu_Test_03_$TAG.ncu-rep > gpurun_out/ncu_full_summary_Test_03_$TAG.txt 2>&1
rm -f gpurun_out/ncu_Test_03_$TAG.ncu-rep
