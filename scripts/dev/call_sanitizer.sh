#!/bin/bash
O=gpurun_out/san; mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(forward_fp32 and small) or (loglik_and_grad_fp32 and small and NORMAL) or device_shuffled or vi_device_steps" > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck.log
tail -15 $O/memcheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "alternative_kernel_paths_agree and FUSED_ENCODE" > $O/memcheck2.log 2>&1; echo "memcheck2 rc=$?" >> $O/memcheck2.log
tail -8 $O/memcheck2.log
