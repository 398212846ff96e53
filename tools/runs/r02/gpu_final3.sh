#!/bin/bash
# Last GPU call of round 2 (4.9 GPU-minutes left): (A) the whole GPU tier on the default library — it carries the
# apply_kernel change and the probe-pass tests — in parallel with the full-decode parity tests on the three
# fd_score epilogue variants (-DFD_EPI=1|2|3, and 2 with -DFD_SIGMOID=1 as lib_epi4); (B) timings, one process per variant.  Every step is bounded by
# what is left of a 262 s budget.
mkdir -p gpurun_out
T0=$(date +%s)
left() { local l=$(( 262 - ($(date +%s) - T0) )); [ $l -lt 1 ] && l=1; echo $l; }
make -s -C oracle all > gpurun_out/g_make.log 2>&1   # once, before five processes would race to do it
python tools/final_ab.py gen C E B > gpurun_out/g_gen.log 2>&1 &
(timeout 200 python -m pytest tests -m gpu -q > gpurun_out/g_pytest.log 2>&1; echo "pytest default rc $?" > gpurun_out/g_pytest.rc) &
for n in 1 2 3 4; do
  (CDAE_B200_LIB=$PWD/cdae_b200/_ab/lib_epi$n.so timeout 150 python -m pytest tests/test_gpu_fulldec.py tests/test_gpu_step_full_size.py -k "fulldec or full_decode" -x -q > gpurun_out/g_epi${n}_pytest.log 2>&1; echo "pytest epi$n rc $?" > gpurun_out/g_epi${n}_pytest.rc) &
done
wait
cat gpurun_out/g_pytest.rc gpurun_out/g_epi*_pytest.rc; tail -2 gpurun_out/g_pytest.log
echo "phase A $(( $(date +%s) - T0 )) s"
for n in 0 1 2 3 4; do
  L=$PWD/cdae_b200/_ab/lib_epi$n.so; [ $n = 0 ] && L=$PWD/cdae_b200/libcdae_b200.so
  CDAE_B200_LIB=$L timeout $(left) python tools/final_ab.py fd C >> gpurun_out/g_fd.jsonl 2>> gpurun_out/g_fd.err
done
echo "fd C $(( $(date +%s) - T0 )) s"
timeout $(left) python tools/final_ab.py topn >> gpurun_out/g_topn.jsonl 2>> gpurun_out/g_topn.err
echo "topn $(( $(date +%s) - T0 )) s"
timeout $(left) python bench.py --config B --no-cpu-baseline --no-extra --no-topn --steps 10 --warmup 3 > gpurun_out/g_bench_B.json 2> gpurun_out/g_bench_B.err
python -c "import json; d=json.load(open('gpurun_out/g_bench_B.json')); print('B value %.2fM e2e %.2fM'%(d['value']/1e6, d['e2e']['value']/1e6), d['kernel_ms_share'])"
echo "bench $(( $(date +%s) - T0 )) s"
for n in 0 1 2 3 4; do
  L=$PWD/cdae_b200/_ab/lib_epi$n.so; [ $n = 0 ] && L=$PWD/cdae_b200/libcdae_b200.so
  CDAE_B200_LIB=$L timeout $(left) python tools/final_ab.py fd E >> gpurun_out/g_fd.jsonl 2>> gpurun_out/g_fd.err
done
echo "fd E $(( $(date +%s) - T0 )) s"
python - <<PY
import json
for f in ("gpurun_out/g_fd.jsonl", "gpurun_out/g_topn.jsonl"):
    try:
        for l in open(f):
            d = json.loads(l)
            if d.get("what") == "fd":
                print(d["shape"], d["lib"].split("/")[-1], "epoch", min(d["epoch_ms"]), {k: round(v, 1) for k, v in d["tflops"].items()})
            else:
                print({k: d[k] for k in d if k not in ("per_class_ms", "lib")})
    except Exception as e:
        print(f, e)
PY
echo "total $(( $(date +%s) - T0 )) s"
