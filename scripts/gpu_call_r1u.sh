#!/bin/bash
# Round-1 final artefact run: parity tests, smoke, bench lines of every workload, reference arm,
# ncu launch list and `ncu --set full` captures (reduced on the box with scripts/ncu_summary.py; the
# summaries are committed under profiles/ -- gpurun brings back at most 64 MiB).
set -x
O=gpurun_out/r1u
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
tail -2 $O/smoke.log
timeout 400 python bench.py > $O/bench_chickenpox_bf16.json 2> $O/bench_chickenpox_bf16.err
timeout 300 python bench.py --workload wind_map_e16 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_wind_bf16.json 2> $O/bench_wind_bf16.err
timeout 300 python bench.py --workload air_quality_map_e8 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_aq_bf16.json 2> $O/bench_aq_bf16.err
timeout 300 python bench.py --workload synthetic_vi_e8 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_vi_bf16.json 2> $O/bench_vi_bf16.err
timeout 300 python bench.py --precision fp32 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_chickenpox_fp32.json 2> $O/bench_chickenpox_fp32.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 60 --csv --log-file $O/launches_chickenpox.csv python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-profile > $O/ncu_launches.log 2>&1
BNF_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -s 60 -c 9 -o $O/ncu_chickenpox_r1u python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile > $O/ncu_cp.log 2>&1
BNF_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm|head_fused' -s 40 -c 12 -o $O/ncu_wind_r1u python bench.py --workload wind_map_e16 --steps 3 --warmup 3 --no-cpu-baseline --no-profile > $O/ncu_wind.log 2>&1
python scripts/ncu_summary.py $O/ncu_chickenpox_r1u.ncu-rep $O/ncu_chickenpox_r1u_summary.csv
python scripts/ncu_summary.py $O/ncu_wind_r1u.ncu-rep $O/ncu_wind_tc_gemm_r1u_summary.csv
rm -f $O/ncu_wind_r1u.ncu-rep
for f in $O/bench_*.json; do echo $f; python - "$f" <<'P'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line); print(d.get('ms_per_step'), d.get('value'), (d.get('e2e') or {}).get('value'), d.get('gpu_launches'), d.get('roofline'), {k:round(v['ms_per_step'],4) for k,v in (d.get('kernels') or {}).items()})
P
done
ls -la $O
