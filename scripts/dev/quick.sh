#!/bin/bash
# Quick GPU check used while iterating on kernels: tensor-core + parity tests, then the
# chickenpox and wind bench lines with their per-kernel CUDA-event times.
#   gpurun --timeout 900 -- 'bash scripts/dev/quick.sh'
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5
for wl in "chickenpox_map_e8 --steps 300 --warmup 20" "wind_map_e16 --steps 5 --warmup 3"; do
timeout 300 python bench.py --workload $wl --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print(d['config']['workload'], round(d['ms_per_step'],4), '%.4g'%d['value'], '%.4g'%d['e2e']['value'], d['gpu_launches'], {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
"
done
