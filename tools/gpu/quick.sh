#!/bin/bash
# quick GPU visit: parity tests, then phase times and a short bench (no ncu)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
if ! grep -q "pytest rc=0" gpurun_out/pytest_gpu.log; then
  timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/sanitizer.log 2>&1; tail -40 gpurun_out/sanitizer.log
  exit 1
fi
timeout 300 python tools/phase_times.py ${1:-Test_03} > gpurun_out/phase.txt 2>&1; cat gpurun_out/phase.txt
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu --e2e-steps 20 > gpurun_out/bench_quick.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench_quick.json; tail -5 gpurun_out/bench.err
