#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f2_pytest.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/f2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f2_smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/f2_smoke.log
timeout 900 python bench.py > gpurun_out/f2_bench.json 2> gpurun_out/f2_bench.err
echo "bench rc $?"
python - <<PY
import json
d=json.load(open("gpurun_out/f2_bench.json"))
print("B value %.2fM e2e %.2fM"%(d["value"]/1e6,d["e2e"]["value"]/1e6), "roofline", d["roofline"]["bound"], round(d["roofline"]["frac"],3), "cpu", round(d["cpu_baseline"]["value"]), round(d["cpu_baseline_multicore"]["value"]))
c=d["configs"]["C"]; print("C value %.2fM e2e %.2fM"%(c["value"]/1e6,c["e2e"]["value"]/1e6), "tensor frac", round(c["roofline"]["frac"],3))
print("topn %.2fM users/s, %.3f ms"%(d["topn"]["users_per_s"]/1e6, d["topn"]["candidate_kernel_ms"]))
PY
