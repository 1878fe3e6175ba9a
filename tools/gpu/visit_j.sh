#!/bin/bash
# group mode of the fused observation kernel: parity tests, then the sweep over environments per CTA
TAG=${1:-r02j}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "plan or group" > gpurun_out/pytest_group_$TAG.log 2>&1; tail -15 gpurun_out/pytest_group_$TAG.log
OUT=gpurun_out/sweep_$TAG.txt; : > $OUT
run() { cfg=$1; shift
  timeout 300 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu --no-sweep --no-configs --no-policy --episodes 0 --profile-steps 10 "$@" 2>> gpurun_out/sweep_$TAG.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1f M  ms/step %.4f  kernels %s  e2e %.1f M (%s, %d chunks, %.0f MB d2h)  plan %s' % (d['value']/1e6, d['ms_per_step'], {k: round(v['ms_per_launch']*v['launches_per_step'],4) for k,v in d['kernels'].items()}, d['e2e']['value']/1e6, d['e2e']['wire'], d['e2e']['chunks'], d['e2e']['d2h_bytes_per_step']/1e6, d['config']['obs_plan']))"; }
for g in 1 7 6 5 4; do echo "== Test_03 group=$g" | tee -a $OUT; FL_OBS_GROUP=$g run Test_03 --e2e-steps 3 | tee -a $OUT; done
for g in 7 6; do echo "== Test_02 parts=0 group=$g" | tee -a $OUT; FL_OBS_PARTS=0 FL_OBS_GROUP=$g run Test_02 --e2e-steps 3 | tee -a $OUT; done
for g in 3 2; do echo "== Test_08 parts=0 group=$g" | tee -a $OUT; FL_OBS_PARTS=0 FL_OBS_GROUP=$g run Test_08 --e2e-steps 3 | tee -a $OUT; done
echo "== Test_08 parts=0 nt=128 group=5" | tee -a $OUT; FL_OBS_PARTS=0 FL_OBS_NT=128 FL_OBS_GROUP=5 run Test_08 --e2e-steps 3 | tee -a $OUT
for cfg in Test_03; do
  echo "== $cfg group=7" >> gpurun_out/phase_${cfg}_$TAG.txt
  FL_OBS_GROUP=7 timeout 600 python tools/phase_times.py $cfg 0 210 >> gpurun_out/phase_${cfg}_$TAG.txt 2>&1; tail -17 gpurun_out/phase_${cfg}_$TAG.txt | grep -v plan; done
