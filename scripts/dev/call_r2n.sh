#!/bin/bash
O=gpurun_out/r2n; mkdir -p $O
run() { env $1 timeout 300 python bench.py $2 --no-cpu-baseline --no-extras 2> $O/bench.err | tee -a $O/bench.jsonl | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$1 $2', round(d['ms_per_step'],4), '%.4g'%d['value'], d['clocks']['sm_mhz'], {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items() if k in ('head_fused','tc_gemm_fwd')})
"; tail -3 $O/bench.err; }
for r in 512 256 128 64 512; do
run "BNF_HEAD_FUSED_MAX_ROWS=$r" "--precision bf16 --workload wind_map_e16 --steps 3 --warmup 3"
done
for r in 512 256 128 64 32; do
run "BNF_HEAD_FUSED_MAX_ROWS=$r" "--precision bf16 --workload air_quality_map_e8 --steps 10 --warmup 3"
done
for r in 512 128 64 32; do
run "BNF_HEAD_FUSED_MAX_ROWS=$r" "--precision bf16x3 --workload air_quality_map_e8 --steps 4 --warmup 3"
done
