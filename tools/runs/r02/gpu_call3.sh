#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_l2.py > gpurun_out/c3_probe.json 2> gpurun_out/c3_probe.err; echo "probe rc $?"; cat gpurun_out/c3_probe.json
for v in u1_mb5 u2_mb5 u2_mb4; do
  export CDAE_B200_LIB=$PWD/cdae_b200/_ab/lib_$v.so
  timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-topn --no-cpu-baseline > gpurun_out/c3_bench_$v.json 2> gpurun_out/c3_bench_$v.err
  echo "$v rc $?"
  python - <<PY
import json
d=json.load(open("gpurun_out/c3_bench_$v.json"))
r=d["roofline"]
print("$v", "value %.2fM e2e %.2fM decode %.1f us frac %.3f"%(d["value"]/1e6,d["e2e"]["value"]/1e6,r["avg_launch_ms"]*1e3,r["frac"]))
PY
done
