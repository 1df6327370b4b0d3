#!/bin/bash
# round-2 call B: first run of the bf16x3 mode (GEMM-level split-operand tests, parity tests), bench timing rework
O=gpurun_out/r2b; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "split_operands" > $O/pytest_split.log 2>&1; tail -15 $O/pytest_split.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > $O/pytest_parity.log 2>&1; tail -40 $O/pytest_parity.log
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q > $O/pytest_tc.log 2>&1; tail -5 $O/pytest_tc.log
for a in "--precision bf16 --steps 20" "--precision bf16 --steps 200" "--precision bf16x3 --steps 20" "--precision fp32 --steps 20"; do
timeout 300 python bench.py $a --no-cpu-baseline 2> $O/bench.err | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$a', round(d['ms_per_step'],4), '%.4g'%d['value'], '%.4g'%d['e2e']['value'], d['gpu_launches'], d['timing'], {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
"
tail -3 $O/bench.err
done
