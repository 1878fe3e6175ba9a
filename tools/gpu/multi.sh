#!/bin/bash
# multi-GPU bench as the driver launches it (torchrun, one rank per GPU), both arms
N=${1:-2}
mkdir -p gpurun_out
python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 1500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; tail -c 300 gpurun_out/bench_ref_n$N.json
