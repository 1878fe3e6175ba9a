#!/bin/bash
TAG=${1:-r02h}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -6 gpurun_out/pytest_gpu_$TAG.log
OUT=gpurun_out/sweep_$TAG.txt; : > $OUT
run() { cfg=$1; shift
  python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu --no-sweep --no-configs --no-policy --episodes 0 --profile-steps 10 "$@" 2>> gpurun_out/sweep_$TAG.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1f M  ms/step %.4f  kernels %s  e2e %.1f M (%s, %d chunks, %.0f MB d2h)  plan %s' % (d['value']/1e6, d['ms_per_step'], {k: round(v['ms_per_launch']*v['launches_per_step'],4) for k,v in d['kernels'].items()}, d['e2e']['value']/1e6, d['e2e']['wire'], d['e2e']['chunks'], d['e2e']['d2h_bytes_per_step']/1e6, d['config']['obs_plan']))"; }
for cfg in Test_03 Test_02 Test_08 Test_14 Test_00; do echo "== $cfg default plan" | tee -a $OUT; run $cfg --e2e-steps 3 | tee -a $OUT; done
echo "== Test_03 flatwalk=0" | tee -a $OUT; FL_OBS_FLATWALK=0 run Test_03 --e2e-steps 3 | tee -a $OUT
echo "== Test_03 no flush" | tee -a $OUT; run Test_03 --e2e-steps 3 --no-flush | tee -a $OUT
echo "== Test_14 no flush" | tee -a $OUT; run Test_14 --e2e-steps 3 --no-flush | tee -a $OUT
echo "== Test_14 parts=2 flatwalk=2" | tee -a $OUT; FL_OBS_FLATWALK=2 FL_OBS_PARTS=2 run Test_14 --e2e-steps 3 | tee -a $OUT
for mode in 2 3; do echo "== Test_03 e2e compact expand=$mode chunks=8" | tee -a $OUT; FL_WIRE_EXPAND=$mode run Test_03 --e2e-wire compact --e2e-chunks 8 --e2e-steps 30 | tee -a $OUT; done
for cfg in Test_14 Test_03; do st=210; [ $cfg = Test_14 ] && st=1400
  echo "== $cfg default" >> gpurun_out/phase_${cfg}_$TAG.txt
  timeout 600 python tools/phase_times.py $cfg 0 $st >> gpurun_out/phase_${cfg}_$TAG.txt 2>&1; tail -17 gpurun_out/phase_${cfg}_$TAG.txt | grep -v plan; done
timeout 600 python tools/reset_bench.py Test_14 0 4 | tee gpurun_out/reset_bench_Test_14_$TAG.json
timeout 300 python tools/policy_bench.py > gpurun_out/policy_bench_$TAG.txt 2>&1; tail -5 gpurun_out/policy_bench_$TAG.txt
