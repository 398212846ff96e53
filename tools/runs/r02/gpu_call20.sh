#!/bin/bash
mkdir -p gpurun_out
for v in default s3_u1_mb4 s4_u1_mb5 s4_u1_mb4; do
  if [ $v = default ]; then unset CDAE_B200_LIB; else export CDAE_B200_LIB=$PWD/cdae_b200/_ab/lib_$v.so; fi
  if [ $v != default ]; then timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "frozen_minibatch or epoch_with_device" > gpurun_out/c20_pytest_$v.log 2>&1; echo "$v pytest rc $?"; tail -1 gpurun_out/c20_pytest_$v.log; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-topn --no-cpu-baseline > gpurun_out/c20_bench_$v.json 2> gpurun_out/c20_bench_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/c20_bench_$v.json"))
r=d["roofline"]
print("$v", "value %.2fM e2e %.2fM decode %.1f us/launch frac %.3f"%(d["value"]/1e6,d["e2e"]["value"]/1e6,r["avg_launch_ms"]*1e3,r["frac"]))
PY
done
