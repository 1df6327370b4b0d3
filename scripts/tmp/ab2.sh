cp build_ab/libB.so bayesnf_b200/libbnf_sm100.so
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -3
for rep in 1 2; do for v in A B; do
cp build_ab/lib$v.so bayesnf_b200/libbnf_sm100.so
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$v cp', round(d['ms_per_step'],4), {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items() if k.startswith('tc_')})
"
timeout 300 python bench.py --workload wind_map_e16 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$v wind', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if k.startswith('tc_')})
"
done; done
