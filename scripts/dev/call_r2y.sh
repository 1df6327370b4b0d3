#!/bin/bash
# r2y: (1) why did the predict block's forward go from 9.3 to 21.3 ms between the r2 and r2x default bench lines:
# per-kernel profile of the forward-only path with the round-2 library (A) and the r2x library (B);
# (2) the rolled TC_FWD_HEAD pass 2 (-DBNF_HEAD_ROLLED, library C): same-box A/B against B, parity tests, ncu summary.
set -x
O=gpurun_out/r2y; mkdir -p $O
cp bayesnf_b200/libbnf_sm100.so /tmp/libB.so
for v in A B A B; do
  if [ $v = A ]; then cp build_ab/libA.so bayesnf_b200/libbnf_sm100.so; else cp /tmp/libB.so bayesnf_b200/libbnf_sm100.so; fi
  TAG=$v PRECS=bf16 timeout 120 python scripts/dev/predict_profile.py >> $O/predict_profile.log 2>&1
done
cat $O/predict_profile.log | grep -v "^+"
Q="--no-cpu-baseline --no-profile --no-extras"
for rep in 1 2; do
  cp /tmp/libB.so bayesnf_b200/libbnf_sm100.so
  timeout 120 python bench.py --steps 20 --warmup 5 $Q > $O/ab_B_bf16_$rep.json 2>> $O/ab.err
  [ $rep = 1 ] && timeout 120 python bench.py --precision bf16x3 --steps 20 --warmup 5 $Q > $O/ab_B_bf16x3_$rep.json 2>> $O/ab.err
  cp build_ab/libC.so bayesnf_b200/libbnf_sm100.so
  timeout 120 python bench.py --steps 20 --warmup 5 $Q > $O/ab_C_bf16_$rep.json 2>> $O/ab.err
  [ $rep = 1 ] && timeout 120 python bench.py --precision bf16x3 --steps 20 --warmup 5 $Q > $O/ab_C_bf16x3_$rep.json 2>> $O/ab.err
done
python - <<'P' | tee $O/ab_summary.txt
import json,glob
for f in sorted(glob.glob('gpurun_out/r2y/ab_*.json')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); print(f.split('/')[-1], d.get('ms_per_step'), d.get('value'), (d.get('e2e') or {}).get('value'), d.get('clocks',{}).get('sm_mhz'))
P
# library C: parity tests + per-kernel summary
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_gpu_C.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_C.log
tail -3 $O/pytest_gpu_C.log
B="--no-cpu-baseline --no-profile --no-extras"
BNF_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none -k regex:'tc_gemm_kernel<256, 3, 7' -s 8 -c 1 -o $O/ncu_fwd_head_bf16_C python bench.py --steps 5 --warmup 3 --repeats 3 $B > $O/n1.log 2>&1
python scripts/ncu_summary.py $O/ncu_fwd_head_bf16_C.ncu-rep $O/ncu_fwd_head_bf16_C_summary.csv
BNF_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none -k regex:'tc_gemm_kernel<256, 3, 7' -s 8 -c 1 -o $O/ncu_fwd_head_bf16x3_C python bench.py --precision bf16x3 --steps 5 --warmup 3 --repeats 3 $B > $O/n2.log 2>&1
python scripts/ncu_summary.py $O/ncu_fwd_head_bf16x3_C.ncu-rep $O/ncu_fwd_head_bf16x3_C_summary.csv
rm -f $O/*.ncu-rep
cp /tmp/libB.so bayesnf_b200/libbnf_sm100.so
ls -la $O
