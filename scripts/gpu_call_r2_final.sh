#!/bin/bash
# Round-2 final artefact run (one B200): parity tests, smoke, bench lines of every workload and mode,
# reference arm, ncu launch lists and `ncu --set full` captures reduced with scripts/ncu_summary.py.
set -x
O=gpurun_out/r2z; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu_r2.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2.log
tail -3 $O/pytest_gpu_r2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r2.log 2>&1; echo "smoke rc=$?" >> $O/smoke_r2.log
tail -3 $O/smoke_r2.log
# the driver's command (headline + extras: bf16x3 record, predict block, wind roofline block)
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default_r2.json 2> $O/bench_default_r2.err
timeout 400 python bench.py --steps 200 --warmup 20 --no-extras > $O/bench_chickenpox_bf16_r2.json 2>> $O/bench.err
timeout 300 python bench.py --precision bf16x3 --steps 50 --warmup 10 --no-extras --no-cpu-baseline > $O/bench_chickenpox_bf16x3_r2.json 2>> $O/bench.err
timeout 300 python bench.py --precision fp32 --steps 20 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_chickenpox_fp32_r2.json 2>> $O/bench.err
for p in bf16 bf16x3; do
timeout 300 python bench.py --precision $p --workload air_quality_map_e8 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_aq_${p}_r2.json 2>> $O/bench.err
timeout 300 python bench.py --precision $p --workload air_quality_mle_zinb_e8 --steps 10 --warmup 4 --no-cpu-baseline > $O/bench_aq_zinb_mle_${p}_r2.json 2>> $O/bench.err
done
timeout 300 python bench.py --workload synthetic_vi_e8 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_vi_bf16_r2.json 2>> $O/bench.err
timeout 300 python bench.py --workload wind_map_e16 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_wind_bf16_r2.json 2>> $O/bench.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference_cpu_r2.json 2>> $O/bench.err
B="--no-cpu-baseline --no-profile --no-extras"
BNF_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 32 --csv --log-file $O/launches_chickenpox_bf16_r2.csv python bench.py --steps 5 --warmup 3 --repeats 3 $B > $O/l1.log 2>&1
BNF_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 32 --csv --log-file $O/launches_chickenpox_bf16x3_r2.csv python bench.py --precision bf16x3 --steps 5 --warmup 3 --repeats 3 $B > $O/l2.log 2>&1
BNF_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -s 64 -c 8 -o $O/ncu_chickenpox_bf16_r2 python bench.py --steps 5 --warmup 3 --repeats 3 $B > $O/n1.log 2>&1
BNF_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -s 64 -c 8 -o $O/ncu_chickenpox_bf16x3_r2 python bench.py --precision bf16x3 --steps 5 --warmup 3 --repeats 3 $B > $O/n2.log 2>&1
# head + activation backward at W = 512: the default kernel and the one-warp-per-row variant (r2v)
BNF_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none -k regex:'head_fused_kernel' -s 3 -c 1 -o $O/ncu_aq_head_fused_r2 python bench.py --workload air_quality_map_e8 --steps 3 --warmup 3 --repeats 3 $B > $O/n3.log 2>&1
BNF_NO_GRAPH=1 BNF_HEAD_ROWS=1 timeout 600 ncu --set full --clock-control none -k regex:'head_rows_kernel' -s 3 -c 1 -o $O/ncu_aq_head_rows_r2 python bench.py --workload air_quality_map_e8 --steps 3 --warmup 3 --repeats 3 $B > $O/n4.log 2>&1
for n in ncu_chickenpox_bf16_r2 ncu_chickenpox_bf16x3_r2 ncu_aq_head_fused_r2 ncu_aq_head_rows_r2; do
  python scripts/ncu_summary.py $O/$n.ncu-rep $O/${n}_summary.csv
done
rm -f $O/*.ncu-rep
for f in $O/bench_*.json; do echo $f; python - "$f" <<'P'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line); print(d.get('ms_per_step'), d.get('value'), (d.get('e2e') or {}).get('value'), d.get('gpu_launches'), d.get('clocks'), d.get('roofline') and {k:d['roofline'][k] for k in ('kernel','bound','frac','achieved')}, {k:round(v['ms_per_step'],4) for k,v in (d.get('kernels') or {}).items()})
P
done
ls -la $O
