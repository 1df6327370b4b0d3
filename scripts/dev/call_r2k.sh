#!/bin/bash
O=gpurun_out/r2k; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; tail -6 $O/pytest.log
run() { env $1 timeout 300 python bench.py $2 --no-cpu-baseline --no-extras 2> $O/bench.err | tee -a $O/bench.jsonl | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$1 $2', round(d['ms_per_step'],4), '%.4g'%d['value'], {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items() if k in ('map_update','encode','tc_fwd_head','tc_fwd_head_x3')})
"; tail -3 $O/bench.err; }
for e in "BNF_UPDATE_BLOCKS_PER_SM=8" "BNF_UPDATE_BLOCKS_PER_SM=4" "BNF_UPDATE_BLOCKS_PER_SM=2" "BNF_UPDATE_BLOCKS_PER_SM=1" "BNF_UPDATE_BLOCKS_PER_SM=8"; do
run "$e" "--precision bf16 --steps 200"
done
for e in "BNF_UPDATE_BLOCKS_PER_SM=8" "BNF_UPDATE_BLOCKS_PER_SM=2"; do
run "$e" "--precision bf16 --workload wind_map_e16 --steps 5 --warmup 3"
done
run "X=0" "--precision bf16x3 --steps 20"
