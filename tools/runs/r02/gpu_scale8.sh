#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/s8_bench_n8.json 2> gpurun_out/s8_bench_n8.err
echo "bench rc $?"
python - <<PY
import json
d=json.load(open("gpurun_out/s8_bench_n8.json"))
print("N=8 B value %.2fM e2e %.2fM"%(d["value"]/1e6,d["e2e"]["value"]/1e6), {k:round(v,3) for k,v in d["kernel_ms_share"].items()}, d["run"]["parallelism"][:90])
for k,c in d.get("configs",{}).items():
    print("  ",k,"value %.2fM e2e %.2fM"%(c["value"]/1e6,c["e2e"]["value"]/1e6), {a:round(b,3) for a,b in c["kernel_ms_share"].items()}, "roofline", c["roofline"]["bound"], round(c["roofline"]["frac"],3))
PY
