#!/bin/bash
# Same-box A/B of library builds: the in-tree library against every tools/gpu/ab/lib*.so (built with -D switches):
# kernel time from bench.py, instruction counters from ncu.   gpurun -- 'bash tools/gpu/ab.sh TAG [notest]'
TAG=${1:-ab}
mkdir -p gpurun_out
LIB=flatland-marl_b200/csrc/libflatland_b200.so
cp $LIB /tmp/libHEAD.so
if [ "$2" != "notest" ]; then
  timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_deep.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
fi
OUT=gpurun_out/${TAG}_ab.txt; : > $OUT
run() { cfg=$1; shift
  timeout 300 python bench.py --config $cfg --steps 40 --warmup 5 --no-cpu --no-sweep --no-configs --no-policy --episodes 0 --profile-steps 20 "$@" 2>> gpurun_out/${TAG}_ab.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1f M  ms/step %.4f  kernels %s  e2e %.1f M (other wire %.1f M)  plan %s' % (d['value']/1e6, d['ms_per_step'], {k: round(v['ms_per_launch']*v['launches_per_step'],4) for k,v in d['kernels'].items()}, d['e2e']['value']/1e6, d['e2e'].get('other_wire',{}).get('value',0)/1e6, d['config']['obs_plan']))"; }
counters() { cfg=$1
  timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'k_observe' -s 6 -c ${CNT:-1} --csv python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu --e2e-steps 1 --profile-steps 1 --episodes 0 --no-sweep --no-configs --no-policy 2>/dev/null |
    python -c "
import sys,csv
rows=[r for r in csv.reader(sys.stdin) if len(r)>10 and r[0].isdigit()]
out={}
for r in rows: out.setdefault(r[4].split('(')[0][-30:],{})[r[-3]]=r[-1]
for k,v in out.items(): print('   ncu', k, ' '.join('%s=%s' % (a.split('.')[0].replace('smsp__','').replace('sm__',''), b) for a,b in v.items()))"; }
for f in /tmp/libHEAD.so tools/gpu/ab/lib*.so; do
  [ -f $f ] || continue
  cp $f $LIB
  echo "== Test_03 lib=$(basename $f .so)" | tee -a $OUT; run Test_03 --e2e-steps 5 | tee -a $OUT; counters Test_03 | tee -a $OUT
  for cfg in $AB_CFGS; do echo "== $cfg lib=$(basename $f .so)" | tee -a $OUT; run $cfg --e2e-steps 5 | tee -a $OUT; done
done
cp /tmp/libHEAD.so $LIB
for l in $AB_LANES; do echo "== Test_03 lib=libHEAD FL_WIRE_LANES=$l" | tee -a $OUT; FL_WIRE_LANES=$l run Test_03 --e2e-steps 40 | tee -a $OUT; done
if [ -n "$AB_BENCH" ]; then
  S0=$SECONDS; timeout 1500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
  echo "bench wall $((SECONDS-S0)) s"; S0=$SECONDS; timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 300 gpurun_out/${TAG}_bench_reference.json
  echo "reference arm wall $((SECONDS-S0)) s"
fi
