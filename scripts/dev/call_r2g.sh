#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; tail -15 $O/pytest.log
for a in "--precision bf16 --steps 200" "--precision bf16x3 --steps 20" "--precision bf16 --workload air_quality_map_e8 --steps 10 --warmup 3"; do
for env in "BNF_NO_BIAS0_WGRAD=0" "BNF_NO_BIAS0_WGRAD=1"; do
env $env BNF_NO_FUSED_ENCODE=1 timeout 300 python bench.py $a --no-cpu-baseline --no-extras 2> $O/bench.err | tee -a $O/bench.jsonl | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$env $a', round(d['ms_per_step'],4), '%.4g'%d['value'], d['gpu_launches'], {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
"
tail -3 $O/bench.err
done; done
