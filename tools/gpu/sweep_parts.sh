#!/bin/bash
# Launch-shape sweep of fl_observe: fused kernel (parts=0) against the split launch with `parts` tree CTAs per environment,
# and the host-buffer step with the full / compact wire format.   gpurun -- 'bash tools/gpu/sweep_parts.sh TAG'
TAG=${1:-sw}
OUT=gpurun_out/sweep_parts_$TAG.txt
: > $OUT
run() {  # config, extra args...
  cfg=$1; shift
  python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu --no-sweep --no-configs --no-policy --episodes 0 --profile-steps 10 "$@" 2>> gpurun_out/sweep_parts_$TAG.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1f M  ms/step %.4f  kernels %s  e2e %.1f M (%s, %d chunks, %.0f MB d2h, %.1f GB/s)  plan %s' % (d['value']/1e6, d['ms_per_step'], {k: round(v['ms_per_launch']*v['launches_per_step'],4) for k,v in d['kernels'].items()}, d['e2e']['value']/1e6, d['e2e']['wire'], d['e2e']['chunks'], d['e2e']['d2h_bytes_per_step']/1e6, d['e2e']['d2h_GBps_per_gpu'], d['config']['obs_plan']))"
}
for cfg in Test_03 Test_02 Test_08 Test_14; do
  for parts in ${PARTS:-0 2 3 4}; do
    echo "== $cfg parts=$parts" | tee -a $OUT
    FL_OBS_PARTS=$parts run $cfg --e2e-steps 3 | tee -a $OUT
  done
done
for wire in full compact; do for ch in 2 4 8 16; do
  echo "== Test_03 e2e wire=$wire chunks=$ch" | tee -a $OUT
  run Test_03 --e2e-wire $wire --e2e-chunks $ch --e2e-steps 30 | tee -a $OUT
done; done
