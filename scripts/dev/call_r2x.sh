#!/bin/bash
# r2x (one call): same-box A/B of the round-2 final library (A) against the warp-uniform / tensor-map
# prefetch / unrolled-step-graph build (B; BNF_GRAPH_UNROLL=1 switches the unrolling off), then the
# full validation + artefact run on B: parity tests, smoke, the driver's bench command, launch list,
# `ncu --set full` summaries in both tensor-core modes, the other workloads.
set -x
O=gpurun_out/r2x; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt
cp bayesnf_b200/libbnf_sm100.so /tmp/libB.so
Q="--no-cpu-baseline --no-profile --no-extras"
for rep in 1 2; do
  cp build_ab/libA.so bayesnf_b200/libbnf_sm100.so
  timeout 120 python bench.py --steps 20 --warmup 5 $Q > $O/ab_A_bf16_$rep.json 2>> $O/ab.err
  cp /tmp/libB.so bayesnf_b200/libbnf_sm100.so
  [ $rep = 1 ] && BNF_GRAPH_UNROLL=1 timeout 120 python bench.py --steps 20 --warmup 5 $Q > $O/ab_B1_bf16_$rep.json 2>> $O/ab.err
  timeout 120 python bench.py --steps 20 --warmup 5 $Q > $O/ab_B8_bf16_$rep.json 2>> $O/ab.err
done
cp build_ab/libA.so bayesnf_b200/libbnf_sm100.so
timeout 120 python bench.py --precision bf16x3 --steps 20 --warmup 5 $Q > $O/ab_A_bf16x3_1.json 2>> $O/ab.err
cp /tmp/libB.so bayesnf_b200/libbnf_sm100.so
timeout 120 python bench.py --precision bf16x3 --steps 20 --warmup 5 $Q > $O/ab_B8_bf16x3_1.json 2>> $O/ab.err
python - <<'P' | tee $O/ab_summary.txt
import json,glob
for f in sorted(glob.glob('gpurun_out/r2x/ab_*.json'))+['gpurun_out/r2x/bench_chickenpox_bf16_r2x.json']:
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f.split('/')[-1], d.get('ms_per_step'), d.get('value'), (d.get('e2e') or {}).get('value'), d.get('clocks',{}).get('sm_mhz'))
P
# ---- validation of B
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_gpu_r2x.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2x.log
tail -4 $O/pytest_gpu_r2x.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r2x.log 2>&1; echo "smoke rc=$?" >> $O/smoke_r2x.log
tail -2 $O/smoke_r2x.log
# the driver's command (headline + extras: bf16x3 record, predict block, wind roofline block)
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default_r2x.json 2> $O/bench_default_r2x.err
B="--no-cpu-baseline --no-profile --no-extras"
BNF_NO_GRAPH=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 32 --csv --log-file $O/launches_chickenpox_bf16_r2x.csv python bench.py --steps 5 --warmup 3 --repeats 3 $B > $O/l1.log 2>&1
BNF_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -s 64 -c 8 -o $O/ncu_chickenpox_bf16_r2x python bench.py --steps 5 --warmup 3 --repeats 3 $B > $O/n1.log 2>&1
python scripts/ncu_summary.py $O/ncu_chickenpox_bf16_r2x.ncu-rep $O/ncu_chickenpox_bf16_r2x_summary.csv
BNF_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none -s 72 -c 9 -o $O/ncu_chickenpox_bf16x3_r2x python bench.py --precision bf16x3 --steps 5 --warmup 3 --repeats 3 $B > $O/n2.log 2>&1
python scripts/ncu_summary.py $O/ncu_chickenpox_bf16x3_r2x.ncu-rep $O/ncu_chickenpox_bf16x3_r2x_summary.csv
rm -f $O/ncu_chickenpox_bf16x3_r2x.ncu-rep
# the other workloads on the final library
timeout 120 python bench.py --steps 200 --warmup 20 $B > $O/bench_chickenpox_bf16_r2x.json 2>> $O/ab.err
timeout 120 python bench.py --precision bf16x3 --steps 50 --warmup 10 $B > $O/bench_chickenpox_bf16x3_r2x.json 2>> $O/bench.err
timeout 150 python bench.py --workload air_quality_mle_zinb_e8 --steps 10 --warmup 4 --no-cpu-baseline > $O/bench_aq_zinb_mle_bf16_r2x.json 2>> $O/bench.err
timeout 150 python bench.py --workload wind_map_e16 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_wind_bf16_r2x.json 2>> $O/bench.err
timeout 150 python bench.py --workload synthetic_vi_e8 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_vi_bf16_r2x.json 2>> $O/bench.err
ls -la $O
