#!/bin/bash
O=gpurun_out/r2j; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; tail -6 $O/pytest.log
run() { env $1 timeout 300 python bench.py $2 --no-cpu-baseline --no-extras 2> $O/bench.err | tee -a $O/bench.jsonl | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$1 $2', round(d['ms_per_step'],4), '%.4g'%d['value'], {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items() if k.startswith('tc_')})
"; tail -3 $O/bench.err; }
for e in "X=0" "BNF_BN_FWD=128" "BNF_BN_DGRAD=128" "BNF_BN_FWD=128 BNF_BN_DGRAD=128" "X=0"; do
run "$e" "--precision bf16 --steps 200"
done
for e in "X=0" "BNF_BN_FWD=128 BNF_BN_DGRAD=128"; do
run "$e" "--precision bf16x3 --steps 20"
done
