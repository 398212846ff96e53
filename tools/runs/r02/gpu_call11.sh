#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_dist.py -x -q > gpurun_out/c11_pytest.log 2>&1
echo "pytest rc $?"; tail -15 gpurun_out/c11_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-extra --allreduce p2p > gpurun_out/c11_bench_n2_p2p.json 2> gpurun_out/c11_bench_n2_p2p.err
python - <<PY
import json
d=json.load(open("gpurun_out/c11_bench_n2_p2p.json"))
print("p2p", "value %.2fM e2e %.2fM"%(d["value"]/1e6,d["e2e"]["value"]/1e6), {k:round(v,3) for k,v in d["kernel_ms_share"].items()})
PY
