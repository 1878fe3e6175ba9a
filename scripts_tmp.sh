for v in "" "FL_OBS_CTAS=12" "FL_OBS_CTAS=10" "FL_OBS_NT=128" "FL_OBS_NT=128 FL_OBS_CTAS=8"; do
  env $v python bench.py --config Test_02 --steps 40 --warmup 5 --no-cpu --e2e-steps 3 --profile-steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v','value %.1fM'%(d['value']/1e6), {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})"
done
python tools/phase_times.py Test_02 2048 100 2>&1 | tail -11
