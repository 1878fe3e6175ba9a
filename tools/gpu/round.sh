#!/bin/bash
# One full GPU-box visit: parity tests, smoke, both bench arms as the driver runs them, ncu launch list, ncu --set full per
# configuration (summaries + raw counters; the reports themselves stay on the box), per-source-line profile of the headline
# kernel, compute-sanitizer memcheck / racecheck of the smoke run.   gpurun -- 'bash tools/gpu/round.sh r02_p [notest]'
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_$TAG.txt 2>&1; nproc >> gpurun_out/gpu_$TAG.txt
if [ "$2" != "notest" ]; then
  timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log; tail -4 gpurun_out/${TAG}_pytest_gpu.log
fi
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
S0=$SECONDS; timeout 1500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
echo "bench wall $((SECONDS-S0)) s"; S0=$SECONDS; timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 400 gpurun_out/${TAG}_bench_reference.json; tail -2 gpurun_out/${TAG}_bench_reference.err
echo "reference arm wall $((SECONDS-S0)) s"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 3 --profile-steps 3 --episodes 0 --no-sweep --no-configs --policy-steps 3 > gpurun_out/${TAG}_ncu_launch.log 2>&1
for cfg in Test_03 Test_02 Test_08 Test_14; do
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_observe' -s 8 -c 2 -o gpurun_out/ncu_${cfg}_$TAG -f \
    python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu --e2e-steps 1 --profile-steps 1 --episodes 0 --no-sweep --no-configs --no-policy > gpurun_out/${TAG}_ncu_full_${cfg}.log 2>&1
  ncu -i gpurun_out/ncu_${cfg}_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_${cfg}_${TAG}_raw.csv 2>/dev/null
  python profiles/summarize.py gpurun_out/ncu_${cfg}_$TAG.ncu-rep > gpurun_out/${TAG}_ncu_full_${cfg}.txt 2>&1
  if [ $cfg = Test_03 ]; then NCU_KRE=k_observe python profiles/hot_lines.py gpurun_out/ncu_${cfg}_$TAG.ncu-rep 'k_observeILi160ELi7ELi0' 70 > gpurun_out/${TAG}_hot_lines_${cfg}.txt 2>&1; fi
  rm -f gpurun_out/ncu_${cfg}_$TAG.ncu-rep
done
timeout 900 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/${TAG}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/${TAG}_sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/${TAG}_sanitizer_racecheck.log
ls -la gpurun_out | grep $TAG
