#!/bin/bash
TAG=${1:-r02g}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -6 gpurun_out/pytest_gpu_$TAG.log
OUT=gpurun_out/sweep_$TAG.txt; : > $OUT
run() { cfg=$1; shift
  python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu --no-sweep --no-configs --no-policy --episodes 0 --profile-steps 10 "$@" 2>> gpurun_out/sweep_$TAG.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1f M  ms/step %.4f  kernels %s  e2e %.1f M (%s, %d chunks, %.0f MB d2h)  plan %s' % (d['value']/1e6, d['ms_per_step'], {k: round(v['ms_per_launch']*v['launches_per_step'],4) for k,v in d['kernels'].items()}, d['e2e']['value']/1e6, d['e2e']['wire'], d['e2e']['chunks'], d['e2e']['d2h_bytes_per_step']/1e6, d['config']['obs_plan']))"; }
for fw in 0 2 3; do
for cfg in Test_03; do echo "== $cfg parts=0 flatwalk=$fw" | tee -a $OUT; FL_OBS_FLATWALK=$fw FL_OBS_PARTS=0 run $cfg --e2e-steps 3 | tee -a $OUT; done
for cfg in Test_14 Test_08 Test_02; do echo "== $cfg parts=2 flatwalk=$fw" | tee -a $OUT; FL_OBS_FLATWALK=$fw FL_OBS_PARTS=2 run $cfg --e2e-steps 3 | tee -a $OUT; done
done
for cfg in Test_14 Test_08; do echo "== $cfg parts=0" | tee -a $OUT; FL_OBS_PARTS=0 run $cfg --e2e-steps 3 | tee -a $OUT; done
for cfg in Test_14 Test_03; do st=210; [ $cfg = Test_14 ] && st=1400
  for parts in 0 2; do echo "== $cfg parts=$parts" >> gpurun_out/phase_${cfg}_$TAG.txt
    FL_OBS_PARTS=$parts timeout 600 python tools/phase_times.py $cfg 0 $st >> gpurun_out/phase_${cfg}_$TAG.txt 2>&1; done; tail -34 gpurun_out/phase_${cfg}_$TAG.txt | grep -v plan; done
