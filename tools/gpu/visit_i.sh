#!/bin/bash
# ncu --set full with source correlation of the fused observation kernel at Test_03: per-source-line stall samples and
# instruction counts (profiles/hot_lines.py, run on the box against the in-tree library), raw metrics; then knob sweeps
TAG=${1:-r02i}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_observe' -s 6 -c 1 -o gpurun_out/ncu_Test_03_$TAG -f \
    python bench.py --config Test_03 --steps 3 --warmup 3 --no-cpu --e2e-steps 1 --profile-steps 1 --episodes 0 --no-sweep --no-configs --no-policy > gpurun_out/ncu_full_Test_03_$TAG.log 2>&1
ncu -i gpurun_out/ncu_Test_03_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_Test_03_${TAG}_raw.csv 2>/dev/null
NCU_KRE=k_observe python profiles/hot_lines.py gpurun_out/ncu_Test_03_$TAG.ncu-rep 'k_observeILi128ELi7ELi0' 70 > gpurun_out/hot_lines_Test_03_$TAG.txt 2>&1
python profiles/summarize.py gpurun_out/ncu_Test_03_$TAG.ncu-rep > gpurun_out/ncu_full_summary_Test_03_$TAG.txt 2>&1
rm -f gpurun_out/ncu_Test_03_$TAG.ncu-rep
OUT=gpurun_out/sweep_$TAG.txt; : > $OUT
run() { cfg=$1; shift
  python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu --no-sweep --no-configs --no-policy --episodes 0 --profile-steps 10 "$@" 2>> gpurun_out/sweep_$TAG.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1f M  ms/step %.4f  kernels %s  e2e %.1f M (%s, %d chunks, %.0f MB d2h)  plan %s' % (d['value']/1e6, d['ms_per_step'], {k: round(v['ms_per_launch']*v['launches_per_step'],4) for k,v in d['kernels'].items()}, d['e2e']['value']/1e6, d['e2e']['wire'], d['e2e']['chunks'], d['e2e']['d2h_bytes_per_step']/1e6, d['config']['obs_plan']))"; }
echo "== Test_02 default plan" | tee -a $OUT; run Test_02 --e2e-steps 3 | tee -a $OUT
for ss in 8 12 16 24; do echo "== Test_03 sortsmall=$ss" | tee -a $OUT; FL_OBS_SORTSMALL=$ss run Test_03 --e2e-steps 3 | tee -a $OUT; done
echo "== Test_03 default" | tee -a $OUT; run Test_03 --e2e-steps 30 | tee -a $OUT
ls -la gpurun_out/*$TAG*
