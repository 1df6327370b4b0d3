timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('cp', d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
"
