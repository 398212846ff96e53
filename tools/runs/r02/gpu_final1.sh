#!/bin/bash
# final 1-GPU validation: the whole GPU test tier, smoke, the default bench line, its ncu launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f1_pytest.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/f1_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f1_smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/f1_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/f1_bench.json 2> gpurun_out/f1_bench.err
echo "bench rc $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f1_bench_ref.json 2> gpurun_out/f1_bench_ref.err
echo "ref rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/f1_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-topn > gpurun_out/f1_b_under_ncu.log 2>&1
echo "ncu launches rc $?"
python - <<PY
import json
d=json.load(open("gpurun_out/f1_bench.json"))
print("B value %.2fM e2e %.2fM"%(d["value"]/1e6,d["e2e"]["value"]/1e6), "roofline", d["roofline"]["bound"], round(d["roofline"]["frac"],3), "cpu", d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline_multicore",{}).get("value"))
c=d["configs"]["C"]; print("C value %.2fM e2e %.2fM"%(c["value"]/1e6,c["e2e"]["value"]/1e6), "tensor frac", round(c["roofline"]["frac"],3), "whole", round(c["roofline"]["whole_step_frac"],3))
print("topn", d["topn"]["users_per_s"], d["topn"]["candidate_kernel_ms"])
r=json.load(open("gpurun_out/f1_bench_ref.json")); print("ref arm", r["value"], r["cpu_baseline"]["build"], r["cpu_baseline"]["sample_users_per_step"])
PY
