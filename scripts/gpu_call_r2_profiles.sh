#!/bin/bash
# Round-2 evidence run: ncu launch lists and `ncu --set full` captures of every kernel class,
# reduced on the box with scripts/ncu_summary.py (gpurun brings back at most 64 MiB).
set -x
O=gpurun_out/r2p; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt
B="--no-cpu-baseline --no-profile --no-extras"
# launch lists (cold-cache, serialised: compare shares)
BNF_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 32 --csv --log-file $O/launches_chickenpox_bf16_r2.csv python bench.py --steps 5 --warmup 3 --repeats 3 $B > $O/l1.log 2>&1
BNF_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 36 --csv --log-file $O/launches_chickenpox_bf16x3_r2.csv python bench.py --precision bf16x3 --steps 5 --warmup 3 --repeats 3 $B > $O/l2.log 2>&1
# full captures: one step of every kernel of the chickenpox step, both tensor-core modes
BNF_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -s 64 -c 8 -o $O/ncu_chickenpox_bf16_r2 python bench.py --steps 5 --warmup 3 --repeats 3 $B > $O/n1.log 2>&1
BNF_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -s 72 -c 9 -o $O/ncu_chickenpox_bf16x3_r2 python bench.py --precision bf16x3 --steps 5 --warmup 3 --repeats 3 $B > $O/n2.log 2>&1
# the update / VI element-wise kernels at the large shapes
BNF_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none -k regex:'map_update|head_fused' -s 4 -c 2 -o $O/ncu_wind_update_r2 python bench.py --workload wind_map_e16 --steps 3 --warmup 3 --repeats 3 $B > $O/n3.log 2>&1
BNF_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none -k regex:'vi_adam|vi_sample|batch_window' -s 6 -c 3 -o $O/ncu_vi_update_r2 python bench.py --workload synthetic_vi_e8 --steps 3 --warmup 3 --repeats 3 $B > $O/n4.log 2>&1
BNF_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none -k regex:'tc_gemm' -s 36 -c 18 -o $O/ncu_wind_tc_gemm_r2 python bench.py --workload wind_map_e16 --steps 3 --warmup 3 --repeats 3 $B > $O/n5.log 2>&1
for n in ncu_chickenpox_bf16_r2 ncu_chickenpox_bf16x3_r2 ncu_wind_update_r2 ncu_vi_update_r2 ncu_wind_tc_gemm_r2; do
  python scripts/ncu_summary.py $O/$n.ncu-rep $O/${n}_summary.csv
done
# keep the two small chickenpox reports (source-level pages), drop the big ones
ls -la $O; rm -f $O/ncu_wind_tc_gemm_r2.ncu-rep $O/ncu_wind_update_r2.ncu-rep $O/ncu_vi_update_r2.ncu-rep
tail -3 $O/*.log
