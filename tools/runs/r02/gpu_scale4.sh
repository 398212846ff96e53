#!/bin/bash
mkdir -p gpurun_out
PROBE_SHAPES=B,D timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 tools/combine_probe.py 2>&1 | grep -E "world" | cut -c1-1500
for mode in p2p nvls; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 --no-extra --allreduce $mode > gpurun_out/s4_bench_n4_$mode.json 2> gpurun_out/s4_bench_n4_$mode.err
python - <<PY
import json
d=json.load(open("gpurun_out/s4_bench_n4_$mode.json"))
print("N=4 $mode value %.2fM e2e %.2fM"%(d["value"]/1e6,d["e2e"]["value"]/1e6), {k:round(v,3) for k,v in d["kernel_ms_share"].items()})
PY
done
