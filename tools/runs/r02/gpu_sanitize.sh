#!/bin/bash
mkdir -p gpurun_out
export CUDA_LAUNCH_BLOCKING=0
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -k "frozen_minibatch or epoch_with_device or bad_csr or q1_scaled or data_loss" > gpurun_out/san_parity.log 2>&1
echo "memcheck parity rc $?"; grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/san_parity.log | tail -5
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_fulldec.py tests/test_gpu_topn_tc.py -x -q > gpurun_out/san_tc.log 2>&1
echo "memcheck tc rc $?"; grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/san_tc.log | tail -5
