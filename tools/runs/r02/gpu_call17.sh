#!/bin/bash
mkdir -p gpurun_out
# trajectory at the 8-GPU weak-scaling size (800K users x 50K items) for the global minibatches 8 x 8192 and 8 x 16384
BQ_U=800000 BQ_I=50000 BQ_EPOCHS=20 BQ_BATCHES="8192,65536,131072" BQ_EVAL_EVERY=2 BQ_ORACLE=0 timeout 900 python tests/experiments/batch_quality.py > gpurun_out/c17_bq_800k.jsonl 2> gpurun_out/c17_bq.err
echo "bq rc $?"; cut -c1-700 gpurun_out/c17_bq_800k.jsonl
for b in 8192 16384 32768; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-topn --no-cpu-baseline --batch-users $b > gpurun_out/c17_bench_b$b.json 2> gpurun_out/c17_bench_b$b.err
  python - <<PY
import json
d=json.load(open("gpurun_out/c17_bench_b$b.json"))
print("batch $b", "value %.2fM e2e %.2fM"%(d["value"]/1e6,d["e2e"]["value"]/1e6), {k:round(v,3) for k,v in d["kernel_ms_share"].items()})
PY
done
