run() { name=$1; shift; env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 --no-topn $EXTRA > gpurun_out/n2_$name.json 2> gpurun_out/n2_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/n2_$name.json")); print("$name", round(d["value"]/1e6,2), "M users/s dev ms", round(d["ms_per_step"],2), "e2e ms", round(d["e2e"]["ms_per_step"],2), {k:round(v,3) for k,v in d["roofline"]["kernel_ms_share"].items()})
except Exception as e: print("$name ERR", e)
PY
}
EXTRA="" run default A=1
EXTRA="" run ring_simple NCCL_ALGO=Ring NCCL_PROTO=Simple
EXTRA="" run skip CDAE_B200_DEBUG_SKIP_ALLREDUCE=1
EXTRA="--batch-users 32768" run b32k A=1
EXTRA="" run ll128 NCCL_PROTO=LL128
