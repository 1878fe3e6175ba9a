#!/bin/bash
# closing check of a round: every GPU test, smoke, both bench arms with the driver's defaults
TAG=${1:-final}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
S0=$SECONDS; timeout 1500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$? wall $((SECONDS-S0)) s"; tail -2 gpurun_out/${TAG}_bench.err
S0=$SECONDS; timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference rc=$? wall $((SECONDS-S0)) s"
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value %.1f M  ms/step %.4f  k_observe %.4f  frac %.4f  e2e %.1f M  caches off %.1f M  episodes %.1f M  sweep %.1f M' % (d['value']/1e6, d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value']/1e6, d['caches']['value_with_caches_off']/1e6, d['episodes']['value']/1e6, d['sweep']['value']/1e6))
for k,v in d['configs'].items(): print(' ', k, '%.1f M' % (v['value']/1e6), 'e2e %.2f M' % (v['e2e']/1e6), v['kernels_ms'])
"
