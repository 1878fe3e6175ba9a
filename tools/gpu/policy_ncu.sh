#!/bin/bash
# ncu evidence for the policy kernels: launch list of one forward + full captures of the main kernels
TAG=${1:-p6}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/policy_launches_$TAG.csv python tools/policy_profile.py Test_03 1024 > gpurun_out/ncu_policy.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_tree_leaf|k_attn_mma' -c 2 -o gpurun_out/prof_policy_a_$TAG -f python tools/policy_profile.py Test_03 1024 >> gpurun_out/ncu_policy.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_tree_p|k_lin' -c 2 -o gpurun_out/prof_policy_b_$TAG -f python tools/policy_profile.py Test_03 1024 >> gpurun_out/ncu_policy.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_lin' -s 10 -c 2 -o gpurun_out/prof_policy_c_$TAG -f python tools/policy_profile.py Test_03 1024 >> gpurun_out/ncu_policy.log 2>&1
tail -3 gpurun_out/ncu_policy.log; ls -la gpurun_out | grep prof_policy
