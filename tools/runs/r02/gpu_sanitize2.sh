#!/bin/bash
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_parity.py -x -q -k "frozen_minibatch or epoch_with_device" > gpurun_out/san_race.log 2>&1
echo "racecheck rc $?"; grep -E "RACECHECK SUMMARY|hazard|passed|failed|ERROR" gpurun_out/san_race.log | tail -6
CDAE_B200_ENCODE=tma timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_parity.py -x -q -k "epoch_with_device" > gpurun_out/san_race_tma.log 2>&1
echo "racecheck tma rc $?"; grep -E "RACECHECK SUMMARY|hazard|passed|failed|ERROR" gpurun_out/san_race_tma.log | tail -4
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_parity.py -x -q -k "frozen_minibatch" > gpurun_out/san_init.log 2>&1
echo "initcheck rc $?"; grep -E "ERROR SUMMARY|Uninitialized|passed|failed" gpurun_out/san_init.log | tail -4
