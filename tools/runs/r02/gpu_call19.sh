#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_topn_tc.py tests/test_gpu_parity.py -x -q -k "topn or golden or recommend" > gpurun_out/c19_pytest.log 2>&1
echo "pytest rc $?"; tail -2 gpurun_out/c19_pytest.log
timeout 600 python -m pytest tests/test_gpu_full_size.py -x -q -k "B or D" > gpurun_out/c19_pytest2.log 2>&1
echo "pytest2 rc $?"; tail -2 gpurun_out/c19_pytest2.log
for v in default tc_epi2; do
  if [ $v = default ]; then unset CDAE_B200_LIB; else export CDAE_B200_LIB=$PWD/cdae_b200/_ab/lib_$v.so; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/c19_bench_$v.json 2> gpurun_out/c19_bench_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/c19_bench_$v.json"))
t=d["topn"]
print("$v topn users/s %.2fM candidate kernel %.3f ms verified %d redone %d frac %.3f"%(t["users_per_s"]/1e6,t["candidate_kernel_ms"],t["verified_users"],t["redone_exact_users"],t["roofline"]["frac"]))
PY
done
