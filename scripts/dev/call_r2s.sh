#!/bin/bash
O=gpurun_out/r2s; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; tail -4 $O/pytest.log
run() { env $1 timeout 600 python bench.py $2 --no-cpu-baseline --no-extras 2> $O/bench.err | tee -a $O/bench.jsonl | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$1 $2', round(d['ms_per_step'],4), 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items() if k.startswith('tc_')})
"; tail -3 $O/bench.err; }
run "X=0" "--precision bf16x3 --steps 20"
run "BNF_BN_FWD=256" "--precision bf16x3 --steps 20"
run "X=0" "--precision bf16x3 --steps 20"
run "BNF_BN_FWD=256" "--precision bf16x3 --steps 20"
