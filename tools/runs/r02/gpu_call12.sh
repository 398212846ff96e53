#!/bin/bash
mkdir -p gpurun_out
for mode in split fused tma; do
  export CDAE_B200_ENCODE=$mode
  timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "epoch_with_device_sampling or q1_scaled" > gpurun_out/c12_pytest_$mode.log 2>&1
  echo "$mode pytest rc $?"; tail -2 gpurun_out/c12_pytest_$mode.log
  timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-topn --no-cpu-baseline > gpurun_out/c12_bench_$mode.json 2> gpurun_out/c12_bench_$mode.err
  python - <<PY
import json,re
d=json.load(open("gpurun_out/c12_bench_$mode.json"))
prof=float(re.search(r"launch, ([0-9.]+) ms per step",d["kernel_share_note"]).group(1))
sh=d["kernel_ms_share"]
print("$mode", "value %.2fM e2e %.2fM"%(d["value"]/1e6,d["e2e"]["value"]/1e6), "per minibatch us:", {k:round(v*prof/13*1e3,1) for k,v in sh.items()})
PY
done
unset CDAE_B200_ENCODE
for mode in split fused tma; do
  CDAE_B200_ENCODE=$mode timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gather_kernel|activate_kernel|encode_fused_kernel" -s 52 -c 2 -o gpurun_out/c12_enc_$mode python tools/profile_run.py --config B > gpurun_out/c12_ncu_$mode.log 2>&1
  echo "ncu $mode rc $?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"decode_kernel|scatter_kernel" -s 52 -c 2 -o gpurun_out/c12_decode python tools/profile_run.py --config B > gpurun_out/c12_ncu_decode.log 2>&1
echo "ncu decode rc $?"
