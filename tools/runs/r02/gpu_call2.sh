#!/bin/bash
# round 2, call 2 (1 GPU): parity of the rewritten decode kernel, then A/B of its unroll / occupancy variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_step_full_size.py -x -q > gpurun_out/c2_pytest.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/c2_pytest.log
for v in default u2_mb4 u4_mb2 u2_mb3; do
  if [ $v = default ]; then unset CDAE_B200_LIB; else export CDAE_B200_LIB=$PWD/cdae_b200/_ab/lib_$v.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-topn --no-cpu-baseline > gpurun_out/c2_bench_$v.json 2> gpurun_out/c2_bench_$v.err
  echo "$v rc $?"
  python - <<PY
import json
d=json.load(open("gpurun_out/c2_bench_$v.json"))
r=d["roofline"]
print("$v", "value %.2fM e2e %.2fM decode %.1f us frac %.3f"%(d["value"]/1e6,d["e2e"]["value"]/1e6,r["avg_launch_ms"]*1e3,r["frac"]), {k:round(v,3) for k,v in d["kernel_ms_share"].items()})
PY
done
