#!/bin/bash
# The very last GPU call of round 2 (3.0 GPU-minutes left): the whole GPU tier on the final default library
# (probe pass ON by default, FD_EPI = 0, the apply_kernel change), then two ncu --set full captures summarised on
# the box — fd_score_kernel at config C's shape (why the epilogue variants of the previous call changed nothing)
# and topn_tc_kernel with probe thresholds — then the default-config bench line with the recommend section.
mkdir -p gpurun_out /tmp/ncu
T0=$(date +%s)
left() { local l=$(( 164 - ($(date +%s) - T0) )); [ $l -lt 1 ] && l=1; echo $l; }
make -s -C oracle all > gpurun_out/h_make.log 2>&1
timeout 100 python -m pytest tests -m gpu -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/h_pytest.log
echo "tests $(( $(date +%s) - T0 )) s"
timeout $(left) ncu --set full --clock-control none -k regex:"fd_score_kernel" -s 1 -c 1 -o /tmp/ncu/fd python tools/profile_run.py --config C --users 18944 --warm 1 --epochs 1 > /tmp/ncu/log_fd.txt 2>&1
python tools/ncu_summary.py /tmp/ncu/fd.ncu-rep > gpurun_out/h_fd_score_ncu_full.txt 2>&1; echo "ncu fd rc $? $(( $(date +%s) - T0 )) s"
timeout $(left) ncu --set full --clock-control none -k regex:"topn_tc_kernel" -s 2 -c 2 -o /tmp/ncu/tn python tools/profile_run.py --config B --warm 2 --epochs 1 --topn > /tmp/ncu/log_tn.txt 2>&1
python tools/ncu_summary.py /tmp/ncu/tn.ncu-rep > gpurun_out/h_topn_probe_ncu_full.txt 2>&1; echo "ncu topn rc $? $(( $(date +%s) - T0 )) s"
timeout $(left) python bench.py --config B --no-cpu-baseline --no-extra --steps 10 --warmup 3 > gpurun_out/h_bench_B.json 2> gpurun_out/h_bench_B.err
python -c "import json; d=json.load(open('gpurun_out/h_bench_B.json')); print('B value %.2fM e2e %.2fM'%(d['value']/1e6, d['e2e']['value']/1e6)); print({k: v for k, v in d['topn'].items() if k != 'roofline'})"
echo "total $(( $(date +%s) - T0 )) s"
