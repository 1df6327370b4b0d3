#!/bin/bash
# same-box A/B of two library builds (build_ab/libA.so = default, libB.so = -DBNF_FWD_EPI12) + tests
cp build_ab/libB.so bayesnf_b200/libbnf_sm100.so
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
cp build_ab/libA.so bayesnf_b200/libbnf_sm100.so
timeout 300 python -m pytest tests -m gpu -q -k "additivity or replays or mini_experiment" 2>&1 | tail -2
for v in A B A B; do
cp build_ab/lib$v.so bayesnf_b200/libbnf_sm100.so
timeout 300 python bench.py --workload wind_map_e16 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$v wind', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k.startswith('tc_')})
"
done
for v in A B; do
cp build_ab/lib$v.so bayesnf_b200/libbnf_sm100.so
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$v cp', round(d['ms_per_step'],4), {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items() if k.startswith('tc_')})
"
timeout 300 python bench.py --workload air_quality_map_e8 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$v aq', round(d['ms_per_step'],4), {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items() if k.startswith('tc_')})
"
done
cp build_ab/libA.so bayesnf_b200/libbnf_sm100.so
