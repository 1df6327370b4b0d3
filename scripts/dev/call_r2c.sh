#!/bin/bash
O=gpurun_out/r2c; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; tail -8 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -5 $O/smoke.log
for a in "--precision bf16 --steps 20" "--precision bf16 --steps 200" "--precision bf16x3 --steps 20" "--precision fp32 --steps 20" "--precision bf16x3 --workload air_quality_map_e8 --steps 10 --warmup 3" "--precision bf16 --workload air_quality_map_e8 --steps 10 --warmup 3"; do
timeout 300 python bench.py $a --no-cpu-baseline 2> $O/bench.err | tee -a $O/bench.jsonl | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$a', round(d['ms_per_step'],4), '%.4g'%d['value'], '%.4g'%d['e2e']['value'], d['gpu_launches'], d['timing'], {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
"
tail -3 $O/bench.err
done
