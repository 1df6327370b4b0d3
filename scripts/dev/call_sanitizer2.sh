#!/bin/bash
O=gpurun_out/san; mkdir -p $O
K="(forward_fp32 and small) or (loglik_and_grad_fp32 and small and ZINB) or device_shuffled or vi_device_steps or map_steps_fp32"
for tool in memcheck racecheck synccheck initcheck; do
timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" > $O/$tool.log 2>&1; echo "$tool rc=$?" >> $O/$tool.log
echo "== $tool"; grep -E "ERROR SUMMARY|passed|failed|rc=" $O/$tool.log | tail -4
done
