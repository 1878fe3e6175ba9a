#!/bin/bash
# One GPU-box visit of round 2: parity tests, smoke, bench (driver arguments), phase profile of every config mid-episode,
# ncu launch list and ncu --set full captures of the observation kernels per config.
#   gpurun -- 'bash tools/gpu/visit.sh TAG [notest] [noncu]'
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
if [ "$2" != "notest" ]; then
  timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
  tail -25 gpurun_out/pytest_gpu_$TAG.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
fi
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
for cfg in Test_03 Test_08 Test_14 Test_02; do
  st=210; [ $cfg = Test_08 ] && st=550; [ $cfg = Test_14 ] && st=1400; [ $cfg = Test_02 ] && st=175
  for parts in ${PARTS:-default}; do
    if [ $parts = default ]; then unset FL_OBS_PARTS; else export FL_OBS_PARTS=$parts; fi
    echo "== $cfg parts=$parts" | tee -a gpurun_out/phase_${cfg}_$TAG.txt
    timeout 600 python tools/phase_times.py $cfg 0 $st >> gpurun_out/phase_${cfg}_$TAG.txt 2>&1; tail -16 gpurun_out/phase_${cfg}_$TAG.txt
  done
  unset FL_OBS_PARTS
done
if [ "$3" != "noncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 3 --profile-steps 3 --episodes 0 --no-sweep --no-configs --no-policy > gpurun_out/ncu_launch_$TAG.log 2>&1
  for cfg in Test_03 Test_08 Test_14 Test_02; do
    timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_observe' -s 3 -c 1 -o gpurun_out/ncu_${cfg}_$TAG -f \
      python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu --e2e-steps 1 --profile-steps 1 --episodes 0 --no-sweep --no-configs --no-policy > gpurun_out/ncu_full_${cfg}_$TAG.log 2>&1
    ncu -i gpurun_out/ncu_${cfg}_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_${cfg}_${TAG}_raw.csv 2>/dev/null
  done
fi
ls -la gpurun_out | head -50
