#!/bin/bash
# policy path on the GPU box: parity tests, timing, ncu launch list of one forward
set -x
TAG=${1:-p1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_policy.py -m gpu -x -q > gpurun_out/pytest_policy.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_policy.log
tail -15 gpurun_out/pytest_policy.log
timeout 600 python tools/policy_bench.py Test_03 1024 30 > gpurun_out/policy_bench_$TAG.json 2> gpurun_out/policy_bench.err; cat gpurun_out/policy_bench_$TAG.json; tail -5 gpurun_out/policy_bench.err
if [ "$2" != "noprof" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_lin|k_tree|k_att|k_prep|k_head|k_choose' -s 400 -c 60 --csv --log-file gpurun_out/policy_launches_$TAG.csv python tools/policy_bench.py Test_03 1024 3 > gpurun_out/ncu_policy.log 2>&1
tail -3 gpurun_out/ncu_policy.log
fi
