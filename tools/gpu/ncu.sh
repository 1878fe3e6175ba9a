#!/bin/bash
# one ncu --set full capture of k_observe in steady state (+ optional launch list)
TAG=${1:-run}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_observe' -s ${2:-120} -c 2 -o gpurun_out/prof_$TAG -f python bench.py --steps 5 --warmup 3 --preroll ${3:-100} --no-cpu --e2e-steps 3 --profile-steps 3 > gpurun_out/ncu_full_bench.log 2>&1
tail -3 gpurun_out/ncu_full_bench.log
ls -la gpurun_out
