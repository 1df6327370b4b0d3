#!/bin/bash
O=gpurun_out/r2i; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; tail -6 $O/pytest.log
for a in "--precision bf16x3 --steps 20" "--precision bf16x3 --steps 20 --workload air_quality_map_e8 --warmup 3" "--precision bf16 --steps 200"; do
timeout 300 python bench.py $a --no-cpu-baseline --no-extras 2> $O/bench.err | tee -a $O/bench.jsonl | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$a', round(d['ms_per_step'],4), '%.4g'%d['value'], d['gpu_launches'], d['roofline'] and (d['roofline']['kernel'], round(d['roofline']['frac'],3)), {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
"
tail -3 $O/bench.err
done
