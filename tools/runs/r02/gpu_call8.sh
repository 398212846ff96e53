#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --extra D > gpurun_out/c10_bench_n8_nvls.json 2> gpurun_out/c10_bench_n8_nvls.err
echo "bench rc $?"
python - <<PY
import json
for f in ("c10_bench_n8_nvls",):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, "value %.2fM e2e %.2fM"%(d["value"]/1e6,d["e2e"]["value"]/1e6), {k:round(v,3) for k,v in d["kernel_ms_share"].items()})
        for k,c in d.get("configs",{}).items():
            print("  ",k,"value %.2fM e2e %.2fM"%(c["value"]/1e6,c["e2e"]["value"]/1e6), {a:round(b,3) for a,b in c["kernel_ms_share"].items()}, "roofline frac", c["roofline"]["frac"] if c.get("roofline") else None)
    except Exception as e: print(f,"parse failed",e)
PY
