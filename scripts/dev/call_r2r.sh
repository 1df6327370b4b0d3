#!/bin/bash
O=gpurun_out/r2r; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; tail -6 $O/pytest.log
run() { env $1 timeout 600 python bench.py $2 --no-cpu-baseline 2> $O/bench.err | tee -a $O/bench.jsonl | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$1 $2', round(d['ms_per_step'],4), 'value %.4g'%d['value'], 'e2e %.4g sync %.4g'%(d['e2e']['value'], d['e2e']['per_step_sync_value']), d['gpu_launches'], 'x3', d.get('parity_mode_bf16x3') and ('%.4g e2e %.4g'%(d['parity_mode_bf16x3']['value'], d['parity_mode_bf16x3']['e2e']['value'])))
"; tail -3 $O/bench.err; }
run "X=0" "--gpus 1 --steps 20 --warmup 5"
run "X=0" "--gpus 1 --steps 20 --warmup 5"
run "X=0" "--steps 200 --no-extras"
run "X=0" "--workload synthetic_vi_e8 --steps 5 --warmup 3 --no-extras"
run "X=0" "--workload air_quality_mle_zinb_e8 --steps 10 --warmup 4 --no-extras"
