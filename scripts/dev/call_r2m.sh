#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; tail -12 $O/pytest.log
run() { env $1 timeout 300 python bench.py $2 --no-cpu-baseline --no-extras 2> $O/bench.err | tee -a $O/bench.jsonl | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$1 $2', round(d['ms_per_step'],4), '%.4g'%d['value'], {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items() if k.startswith('tc_')})
"; tail -3 $O/bench.err; }
for e in "BNF_X3_WGRAD_SEGMENTED=0" "BNF_X3_WGRAD_SEGMENTED=1" "BNF_X3_WGRAD_SEGMENTED=0"; do
run "$e" "--precision bf16x3 --steps 20"
done
for e in "BNF_X3_WGRAD_SEGMENTED=0" "BNF_X3_WGRAD_SEGMENTED=1"; do
run "$e" "--precision bf16x3 --steps 10 --workload air_quality_map_e8 --warmup 3"
done
