set -x
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5
for env in "X=1"; do
env $env timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$env', d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
"
done
timeout 300 python bench.py --workload wind_map_e16 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('wind', d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
"
