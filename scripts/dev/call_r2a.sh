#!/bin/bash
# round-2 call A: tcgen05 accumulation rounding experiment + EPI16 A/B
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt
timeout 300 python scripts/dev/acc_rounding.py > $O/acc.log 2>&1; tail -40 $O/acc.log
timeout 1500 bash scripts/dev/ab_epi16.sh > $O/epi16.log 2>&1; cat $O/epi16.log
