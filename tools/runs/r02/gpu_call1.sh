#!/bin/bash
# round 2, call 1 (1 GPU): bench (B + C + topn + CPU baselines), ncu launch list + full captures, GPU tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
nproc > gpurun_out/c1_nproc.txt; lscpu | head -20 >> gpurun_out/c1_nproc.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
echo "bench rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/c1_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-topn > gpurun_out/c1_b_under_ncu.log 2>&1
echo "ncu launches rc $?"
# one minibatch of config B = gather, activate, decode, hidden_backward, scatter, apply: skip the 2 warm epochs (13 minibatches each)
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"gather_kernel|activate_kernel|decode_kernel|hidden_backward_kernel|scatter_kernel|apply_kernel" -s 156 -c 6 \
    -o gpurun_out/c1_train python tools/profile_run.py --config B > gpurun_out/c1_ncu_train.log 2>&1
echo "ncu train rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"topn_tc_kernel" -s 1 -c 1 \
    -o gpurun_out/c1_topn python tools/profile_run.py --config B --warm 1 --epochs 0 --topn > gpurun_out/c1_ncu_topn.log 2>&1
echo "ncu topn rc $?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1
echo "pytest rc $?"
tail -5 gpurun_out/c1_pytest.log
head -c 1500 gpurun_out/c1_bench.json
