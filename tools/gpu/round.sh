#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list, ncu full capture of the top kernels.
set -x
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
if ! grep -q "pytest rc=0" gpurun_out/pytest_gpu.log; then
  timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/sanitizer.log 2>&1; tail -40 gpurun_out/sanitizer.log
  exit 1
fi
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench.err
if [ "$2" != "noprof" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 5 --warmup 3 --preroll 40 --no-cpu --e2e-steps 3 --profile-steps 3 > gpurun_out/ncu_launch_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_observe|k_step' -s 90 -c 4 -o gpurun_out/prof_$TAG -f python bench.py --steps 5 --warmup 3 --preroll 40 --no-cpu --e2e-steps 3 --profile-steps 3 > gpurun_out/ncu_full_bench.log 2>&1
fi
ls -la gpurun_out
