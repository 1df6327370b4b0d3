#!/bin/bash
# Round-1 session-2 GPU call: parity tests, A/B of the fused step pieces, wind shard, launch list.
set -x
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -8 gpurun_out/pytest.log
python scripts/dev/diag1.py 2>&1 | tail -12
timeout 300 python bench.py --steps 300 --warmup 20 > gpurun_out/bench_cp_default.json 2> gpurun_out/bench_cp_default.err
BNF_PDL=0 timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/bench_cp_nopdl.json 2> gpurun_out/bench_cp_nopdl.err
BNF_LEGACY_STEP=1 timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/bench_cp_legacy.json 2> gpurun_out/bench_cp_legacy.err
timeout 300 python bench.py --workload wind_map_e16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wind.json 2> gpurun_out/bench_wind.err
BNF_LEGACY_STEP=1 timeout 300 python bench.py --workload wind_map_e16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wind_legacy.json 2> gpurun_out/bench_wind_legacy.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 60 --csv --log-file gpurun_out/launches_chickenpox.csv python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/ncu_bench.log 2>&1
for f in gpurun_out/bench_*.json; do echo $f; python - "$f" <<'P'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
P
done
