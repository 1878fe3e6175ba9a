#!/bin/bash
# multi-GPU bench as the driver launches it (torchrun, one rank per GPU), both arms.   gpurun --gpus N -- 'bash tools/gpu/multi.sh N TAG'
N=${1:-2}; TAG=${2:-multi}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
python -c "
import json,sys
d=json.load(open('gpurun_out/${TAG}_bench_n$N.json'))
print('n_gpus',d['n_gpus'],'value %.1f M'%(d['value']/1e6),'scaling',d['scaling'],'e2e %.1f M'%(d['e2e']['value']/1e6))
s=d['sweep']; print('sweep value %.1f M'%(s['value']/1e6),'scaling',s['scaling'],'ms/step',round(s['ms_per_step'],4),'e2e %.1f M'%(s['e2e']['value']/1e6), s['config']['workload'][:90])
"; tail -3 gpurun_out/${TAG}_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_reference_n$N.json 2> gpurun_out/${TAG}_bench_reference_n$N.err; tail -c 300 gpurun_out/${TAG}_bench_reference_n$N.json
