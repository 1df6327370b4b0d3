#!/bin/bash
O=gpurun_out/r2q; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; tail -12 $O/pytest.log
run() { env $1 timeout 600 python bench.py $2 --no-cpu-baseline --no-extras 2> $O/bench.err | tee -a $O/bench.jsonl | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$1 $2', round(d['ms_per_step'],4), '%.4g'%d['value'], d['clocks']['sm_mhz'], {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items() if k.startswith('tc_')})
"; tail -3 $O/bench.err; }
run "X=0" "--precision bf16 --steps 200"
run "X=0" "--precision bf16x3 --steps 20"
run "X=0" "--precision bf16 --workload wind_map_e16 --steps 3 --warmup 3"
run "X=0" "--precision bf16x3 --workload wind_map_e16 --steps 3 --warmup 3"
run "X=0" "--precision bf16 --workload air_quality_map_e8 --steps 10 --warmup 3"
run "X=0" "--precision bf16x3 --workload air_quality_map_e8 --steps 4 --warmup 3"
