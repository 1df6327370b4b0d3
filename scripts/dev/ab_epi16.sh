#!/bin/bash
# Round-2 first GPU call: is the sixteen-warp / 16-column-half epilogue variant (-DBNF_EPI16) correct
# and faster?  Build both libraries HERE (CPU container) first:
#   F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC --threads 4"
#   mkdir -p build_ab; nvcc $F -o build_ab/libA.so bayesnf_b200/csrc/*.cu
#   nvcc $F -DBNF_EPI16 -o build_ab/libB.so bayesnf_b200/csrc/*.cu
#   gpurun --timeout 900 -- 'bash scripts/dev/ab_epi16.sh'
# (remove build_ab/ afterwards: it travels with every gpurun snapshot)
cp build_ab/libB.so bayesnf_b200/libbnf_sm100.so
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -6
for v in A B A B; do
cp build_ab/lib$v.so bayesnf_b200/libbnf_sm100.so
for wl in "chickenpox_map_e8 --steps 300 --warmup 20" "wind_map_e16 --steps 5 --warmup 3" "air_quality_map_e8 --steps 20 --warmup 3"; do
timeout 300 python bench.py --workload $wl --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$v', d['config']['workload'], round(d['ms_per_step'],4), {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items() if k.startswith('tc_')})
"
done; done
cp build_ab/libA.so bayesnf_b200/libbnf_sm100.so
