#!/bin/bash
# head_rows_kernel (one warp per row) vs head_fused_kernel: tests in both settings, then A/B
O=gpurun_out/r2u; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --tb=short > $O/pytest.log 2>&1; tail -n 25 $O/pytest.log
BNF_HEAD_ROWS=0 timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q --no-header > $O/pytest_old.log 2>&1; tail -n 3 $O/pytest_old.log
run() { env $1 timeout 900 python bench.py $2 --no-cpu-baseline --no-extras 2> $O/bench.err | tee -a $O/bench.jsonl | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$1 $2', round(d['ms_per_step'],4), 'value %.4g'%d['value'], {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items() if k.startswith('head')}, d['clocks']['sm_mhz'])
"; tail -n 2 $O/bench.err; }
for w in "air_quality_map_e8 --steps 10" "air_quality_mle_zinb_e8 --steps 10 --warmup 4" ; do
  for p in bf16 bf16x3; do
    run "X=0" "--workload $w --precision $p"
    run "BNF_HEAD_ROWS=0" "--workload $w --precision $p"
  done
done
run "X=0" "--workload wind_map_e16 --precision bf16 --steps 4 --warmup 3"
run "BNF_HEAD_ROWS=0" "--workload wind_map_e16 --precision bf16 --steps 4 --warmup 3"
run "BNF_HEAD_ROWS_MAX=256" "--workload air_quality_map_e8 --steps 10 --precision bf16"
run "BNF_HEAD_ROWS_MAX=128" "--workload air_quality_map_e8 --steps 10 --precision bf16"
