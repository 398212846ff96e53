#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fulldec.py tests/test_gpu_step_full_size.py -x -q > gpurun_out/c22_pytest.log 2>&1
echo "pytest rc $?"; tail -2 gpurun_out/c22_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-topn --no-cpu-baseline > gpurun_out/c22_bench.json 2> gpurun_out/c22_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/c22_bench.json"))
print("value %.2fM e2e %.2fM (ms %.3f vs %.3f)"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],d["e2e"]["ms_per_step"]))
PY
