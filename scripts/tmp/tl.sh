cp build_ab/libX.so bayesnf_b200/libbnf_sm100.so
for m in 7 5 0; do
BNF_NO_GRAPH=1 BNF_TC_TL=$m timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-profile 2>&1 | grep "^TL" | head -40
done
