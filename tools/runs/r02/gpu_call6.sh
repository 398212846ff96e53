#!/bin/bash
# round 2, call 6 (8 GPUs): parity at world 8, then the default 8-GPU bench line (B + configs D, E) and B with NCCL
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/dist_worker.py > gpurun_out/c6_dist.log 2>&1
echo "dist rc $?"; grep -E "FAIL|dist_worker" gpurun_out/c6_dist.log | head -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/c6_bench_n8.json 2> gpurun_out/c6_bench_n8.err
echo "bench rc $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-extra --allreduce nccl > gpurun_out/c6_bench_n8_nccl.json 2> gpurun_out/c6_bench_n8_nccl.err
echo "nccl rc $?"
python - <<PY
import json
for f in ("c6_bench_n8","c6_bench_n8_nccl"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, "value %.2fM e2e %.2fM"%(d["value"]/1e6,d["e2e"]["value"]/1e6), {k:round(v,3) for k,v in d["kernel_ms_share"].items()})
        for k,c in d.get("configs",{}).items():
            print("  ",k,"value %.2fM e2e %.2fM"%(c["value"]/1e6,c["e2e"]["value"]/1e6), {a:round(b,3) for a,b in c["kernel_ms_share"].items()}, "roofline frac", c["roofline"]["frac"] if c.get("roofline") else None)
    except Exception as e: print(f,"parse failed",e)
PY
tail -3 gpurun_out/c6_bench_n8.err
