set -x
mkdir -p gpurun_out
# one full-batch MAP step of the chickenpox workload = 10 kernels; skip the first steps, capture 2 steps
BNF_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on  -s 60 -c 18 -o gpurun_out/ncu_chickenpox_r1t python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/ncu_cp.log 2>&1
ls -la gpurun_out/*.ncu-rep
