#!/bin/bash
# round 2, call 4 (2 GPUs): multi-GPU parity (NCCL and the fused peer-memory combine step), then bench B at N=2 both ways
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_worker.py > gpurun_out/c4_dist.log 2>&1
echo "dist rc $?"; grep -E "FAIL|dist_worker" gpurun_out/c4_dist.log | head -20
for mode in nccl p2p; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-extra --allreduce $mode > gpurun_out/c4_bench_n2_$mode.json 2> gpurun_out/c4_bench_n2_$mode.err
  echo "$mode rc $?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c4_bench_n2_$mode.json"))
    print("$mode", "value %.2fM e2e %.2fM"%(d["value"]/1e6,d["e2e"]["value"]/1e6), {k:round(v,3) for k,v in d["kernel_ms_share"].items()}, d["run"]["parallelism"])
except Exception as e: print("parse failed", e)
PY
done
tail -5 gpurun_out/c4_bench_n2_p2p.err
