#!/bin/bash
# racecheck reports hazards inside `tcgen05.alloc.cta_group::2` (bnf_tc.cu: the instruction itself is both
# accesses) in every CTA-pair kernel: is that the tool's model of the instruction or the r2x/r2y changes?
# Same test under racecheck with the round-2 library (A, before those changes) and the final one.
set -x
O=gpurun_out/r2z2; mkdir -p $O
cp bayesnf_b200/libbnf_sm100.so /tmp/libF.so
for v in A F; do
  if [ $v = A ]; then cp build_ab/libA.so bayesnf_b200/libbnf_sm100.so; else cp /tmp/libF.so bayesnf_b200/libbnf_sm100.so; fi
  timeout 100 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "fused_head_matches_two_kernel_path" > $O/racecheck_$v.log 2>&1; echo "racecheck $v rc=$?" >> $O/racecheck_$v.log
  grep -E "Error: Race|RACECHECK SUMMARY|passed|failed|rc=" $O/racecheck_$v.log | cut -c1-220 | tail -8
done
cp /tmp/libF.so bayesnf_b200/libbnf_sm100.so
BNF_CTA2=0 timeout 100 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "fused_head_matches_two_kernel_path or (bf16_tc_vs_oracle and NORMAL) or (edge_shapes_against_oracle and bf16x3 and no_pad_column and 257)" > $O/racecheck_F_cta1.log 2>&1; echo "racecheck F (single-CTA tiles) rc=$?" >> $O/racecheck_F_cta1.log
grep -E "Error: Race|RACECHECK SUMMARY|passed|failed|rc=" $O/racecheck_F_cta1.log | cut -c1-220 | tail -8
