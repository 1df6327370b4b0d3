#!/bin/bash
# r2z: final validation + artefact run of the round-2 library (rolled TC_FWD_HEAD pass 2, warp-uniform
# operands, tensor-map prefetch, unrolled step graph, 65 536-row forecast slabs): parity tests, smoke,
# the driver's bench command, per-mode bench lines, launch list, ncu summaries, the other workloads,
# compute-sanitizer memcheck / racecheck over the tests that run the changed TC_FWD_HEAD epilogues.
set -x
O=gpurun_out/r2z; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_gpu_r2z.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2z.log
tail -4 $O/pytest_gpu_r2z.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r2z.log 2>&1; echo "smoke rc=$?" >> $O/smoke_r2z.log
tail -2 $O/smoke_r2z.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default_r2z.json 2> $O/bench_default_r2z.err
B="--no-cpu-baseline --no-profile --no-extras"
timeout 120 python bench.py --steps 200 --warmup 20 $B > $O/bench_chickenpox_bf16_r2z.json 2>> $O/bench.err
timeout 120 python bench.py --precision bf16x3 --steps 50 --warmup 10 $B > $O/bench_chickenpox_bf16x3_r2z.json 2>> $O/bench.err
BNF_NO_GRAPH=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 32 --csv --log-file $O/launches_chickenpox_bf16_r2z.csv python bench.py --steps 5 --warmup 3 --repeats 3 $B > $O/l1.log 2>&1
BNF_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none -s 64 -c 8 -o $O/ncu_chickenpox_bf16_r2z python bench.py --steps 5 --warmup 3 --repeats 3 $B > $O/n1.log 2>&1
python scripts/ncu_summary.py $O/ncu_chickenpox_bf16_r2z.ncu-rep $O/ncu_chickenpox_bf16_r2z_summary.csv
BNF_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none -s 72 -c 9 -o $O/ncu_chickenpox_bf16x3_r2z python bench.py --precision bf16x3 --steps 5 --warmup 3 --repeats 3 $B > $O/n2.log 2>&1
python scripts/ncu_summary.py $O/ncu_chickenpox_bf16x3_r2z.ncu-rep $O/ncu_chickenpox_bf16x3_r2z_summary.csv
for n in ncu_chickenpox_bf16_r2z ncu_chickenpox_bf16x3_r2z; do
  ncu -i $O/$n.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
st=[c for c in h if c.startswith('smsp__pcsamp_warps_issue_stalled_') and not c.endswith('_not_issued')]
for r in rows[2:]:
    tot=sum(float(r[h.index(c)] or 0) for c in st) or 1
    v=sorted([(float(r[h.index(c)] or 0)/tot*100, c.replace('smsp__pcsamp_warps_issue_stalled_','')) for c in st], reverse=True)[:8]
    k=r[h.index('Kernel Name')]; k=k[:k.index('(')] if '(' in k else k
    print('%-48s %7.2f us  '%(k[:48], float(r[h.index('gpu__time_duration.sum')])) + '  '.join('%s %.0f%%'%(n,x) for x,n in v))
" > $O/stalls_$n.txt
done
rm -f $O/*.ncu-rep
timeout 150 python bench.py --workload air_quality_mle_zinb_e8 --steps 10 --warmup 4 --no-cpu-baseline > $O/bench_aq_zinb_mle_bf16_r2z.json 2>> $O/bench.err
timeout 150 python bench.py --workload wind_map_e16 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_wind_bf16_r2z.json 2>> $O/bench.err
timeout 150 python bench.py --workload synthetic_vi_e8 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_vi_bf16_r2z.json 2>> $O/bench.err
K="fused_head_matches_two_kernel_path or (bf16_tc_vs_oracle and NORMAL) or (edge_shapes_against_oracle and bf16x3 and no_pad_column and 257)"
for tool in memcheck racecheck; do
  timeout 110 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "$K" > $O/$tool.log 2>&1; echo "$tool rc=$?" >> $O/$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" $O/$tool.log | tail -4
done
ls -la $O
