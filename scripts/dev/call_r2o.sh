#!/bin/bash
O=gpurun_out/r2o; mkdir -p $O
timeout 600 python bench.py --precision bf16x3 --workload wind_map_e16 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_wind_bf16x3_r2.json 2> $O/bench.err; tail -3 $O/bench.err
python - <<'P'
import json
for line in open('gpurun_out/r2o/bench_wind_bf16x3_r2.json'):
    if line.startswith('{'):
        d=json.loads(line); print(round(d['ms_per_step'],3), '%.4g'%d['value'], d['clocks'], d['roofline'] and {k:d['roofline'][k] for k in ('kernel','bound','frac','achieved')}, {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
P
BNF_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none -k regex:'tc_gemm' -s 12 -c 6 -o $O/ncu_wind_x3 python bench.py --precision bf16x3 --workload wind_map_e16 --steps 1 --warmup 3 --repeats 3 --no-cpu-baseline --no-profile --no-extras > $O/n.log 2>&1
python scripts/ncu_summary.py $O/ncu_wind_x3.ncu-rep $O/ncu_wind_tc_gemm_bf16x3_r2_summary.csv; rm -f $O/ncu_wind_x3.ncu-rep; tail -2 $O/n.log
