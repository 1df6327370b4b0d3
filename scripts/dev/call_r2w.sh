#!/bin/bash
# r2w: (1) parity tests at HEAD, (2) `ncu --set full` reports WITH source pages of the chickenpox step in
# both tensor-core modes (kept: read offline for warp-stall attribution), (3) experiment build:
# epilogue ablation masks and per-tile / per-chunk clock64 timelines at the chickenpox shape.
set -x
O=gpurun_out/r2w; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
B="--no-cpu-baseline --no-profile --no-extras"
BNF_NO_GRAPH=1 timeout 400 ncu --set full --clock-control none --import-source on -s 64 -c 8 -o $O/ncu_chickenpox_bf16_r2w python bench.py --steps 5 --warmup 3 --repeats 3 $B > $O/n1.log 2>&1
BNF_NO_GRAPH=1 timeout 400 ncu --set full --clock-control none --import-source on -s 72 -c 9 -o $O/ncu_chickenpox_bf16x3_r2w python bench.py --precision bf16x3 --steps 5 --warmup 3 --repeats 3 $B > $O/n2.log 2>&1
ls -la $O
cp bayesnf_b200/libbnf_sm100.so /tmp/lib_default.so
cp build_ab/libX.so bayesnf_b200/libbnf_sm100.so
export BNF_NO_GRAPH=1 WL=chickenpox_map_e8 ALL_KERNELS=1
TAG=abl MASKS=0,4,8,12,0 timeout 200 python scripts/epi_experiment.py > $O/ablation_bf16.log 2>&1
for mode in 0 5 7; do
  TAG=tl$mode MASKS=0,0 BNF_TC_TL=$mode timeout 200 python scripts/epi_experiment.py > $O/tl_bf16_mode$mode.log 2>&1
done
for mode in 7; do
  PREC=bf16x3 TAG=tlx3_$mode MASKS=0,0 BNF_TC_TL=$mode timeout 200 python scripts/epi_experiment.py > $O/tl_bf16x3_mode$mode.log 2>&1
done
cp /tmp/lib_default.so bayesnf_b200/libbnf_sm100.so
tail -4 $O/ablation_bf16.log
ls -la $O
