#!/bin/bash
# quick visit: parity tests of the observation plans, then Test_03 / Test_02 / Test_08 / Test_14 at the default plan (+ phase times)
TAG=${1:-r02k}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_deep.py -m gpu -q -x > gpurun_out/pytest_$TAG.log 2>&1; tail -5 gpurun_out/pytest_$TAG.log
OUT=gpurun_out/sweep_$TAG.txt; : > $OUT
run() { cfg=$1; shift
  timeout 300 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu --no-sweep --no-configs --no-policy --episodes 0 --profile-steps 10 "$@" 2>> gpurun_out/sweep_$TAG.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1f M  ms/step %.4f  kernels %s  e2e %.1f M (%s, %d chunks, %.0f MB d2h)  plan %s' % (d['value']/1e6, d['ms_per_step'], {k: round(v['ms_per_launch']*v['launches_per_step'],4) for k,v in d['kernels'].items()}, d['e2e']['value']/1e6, d['e2e']['wire'], d['e2e']['chunks'], d['e2e']['d2h_bytes_per_step']/1e6, d['config']['obs_plan']))"; }
for cfg in ${CFGS:-Test_03 Test_02 Test_08 Test_14}; do echo "== $cfg default" | tee -a $OUT; run $cfg --e2e-steps 3 | tee -a $OUT; done
for x in $EXTRA; do echo "== Test_03 $x" | tee -a $OUT; env $x bash -c "$(declare -f run); TAG=$TAG; run Test_03 --e2e-steps 3" | tee -a $OUT; done
echo "== Test_03 default" >> gpurun_out/phase_Test_03_$TAG.txt
timeout 600 python tools/phase_times.py Test_03 0 210 >> gpurun_out/phase_Test_03_$TAG.txt 2>&1; tail -17 gpurun_out/phase_Test_03_$TAG.txt | grep -v plan
