mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for c in 1 2 4 8 16; do
  python bench.py --steps 30 --warmup 5 --no-cpu --e2e-steps 40 --e2e-chunks $c --profile-steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chunks',d['e2e']['chunks'],'e2e %.2fM'%(d['e2e']['value']/1e6),'value %.1fM'%(d['value']/1e6))"
done
