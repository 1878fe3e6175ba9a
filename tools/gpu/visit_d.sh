#!/bin/bash
# visit d: tests, phase profiles after the sort rewrite, fused 128x7, e2e compact timing breakdown
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
OUT=gpurun_out/sweep_$TAG.txt; : > $OUT
run() { cfg=$1; shift
  python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu --no-sweep --no-configs --no-policy --episodes 0 --profile-steps 10 "$@" 2>> gpurun_out/sweep_$TAG.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1f M  ms/step %.4f  kernels %s  e2e %.1f M (%s, %d chunks, %.0f MB d2h)  plan %s' % (d['value']/1e6, d['ms_per_step'], {k: round(v['ms_per_launch']*v['launches_per_step'],4) for k,v in d['kernels'].items()}, d['e2e']['value']/1e6, d['e2e']['wire'], d['e2e']['chunks'], d['e2e']['d2h_bytes_per_step']/1e6, d['config']['obs_plan']))"; }
echo "== Test_03 fused" | tee -a $OUT; FL_OBS_PARTS=0 run Test_03 --e2e-steps 3 | tee -a $OUT
for cfg in Test_14 Test_08; do for parts in 0 2; do
  echo "== $cfg parts=$parts" | tee -a $OUT; FL_OBS_PARTS=$parts run $cfg --e2e-steps 3 | tee -a $OUT
done; done
echo "== Test_14 parts=2 bmglobal=1 treent=512" | tee -a $OUT; FL_OBS_PARTS=2 FL_OBS_BMGLOBAL=1 FL_OBS_TREENT=512 run Test_14 --e2e-steps 3 | tee -a $OUT
echo "== Test_14 parts=4 bmglobal=1 treent=512" | tee -a $OUT; FL_OBS_PARTS=4 FL_OBS_BMGLOBAL=1 FL_OBS_TREENT=512 run Test_14 --e2e-steps 3 | tee -a $OUT
echo "== Test_14 parts=4 bmglobal=1 treent=256" | tee -a $OUT; FL_OBS_PARTS=4 FL_OBS_BMGLOBAL=1 FL_OBS_TREENT=256 run Test_14 --e2e-steps 3 | tee -a $OUT
echo "== Test_08 parts=2 treent=128" | tee -a $OUT; FL_OBS_PARTS=2 FL_OBS_TREENT=128 run Test_08 --e2e-steps 3 | tee -a $OUT
for cfg in Test_14 Test_08; do st=550; [ $cfg = Test_14 ] && st=1400
  for parts in 0 2; do echo "== $cfg parts=$parts" >> gpurun_out/phase_${cfg}_$TAG.txt
    FL_OBS_PARTS=$parts timeout 600 python tools/phase_times.py $cfg 0 $st >> gpurun_out/phase_${cfg}_$TAG.txt 2>&1; done; tail -34 gpurun_out/phase_${cfg}_$TAG.txt | grep -v plan; done
for mode in 0 1 2; do for ch in 4 8; do
  echo "== Test_03 e2e compact expand=$mode chunks=$ch" | tee -a $OUT
  FL_WIRE_EXPAND=$mode FL_WIRE_TIMING=1 run Test_03 --e2e-wire compact --e2e-chunks $ch --e2e-steps 20 2> gpurun_out/wire_timing_${mode}_$ch.txt | tee -a $OUT
  tail -3 gpurun_out/wire_timing_${mode}_$ch.txt | tee -a $OUT
done; done
nproc; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core" 
