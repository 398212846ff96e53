#!/bin/bash
# ncu --set full captures summarised ON THE BOX (the reports are too large to bring back)
mkdir -p gpurun_out /tmp/ncu
for mode in split fused tma; do
  CDAE_B200_ENCODE=$mode timeout 600 ncu --set full --clock-control none -k regex:"gather_kernel|activate_kernel|encode_fused_kernel" -s 52 -c 2 -o /tmp/ncu/enc_$mode python tools/profile_run.py --config B > /tmp/ncu/log_$mode.txt 2>&1
  python tools/ncu_summary.py /tmp/ncu/enc_$mode.ncu-rep > gpurun_out/c13_encode_${mode}_ncu_full.txt
  echo "ncu $mode rc $?"
done
timeout 600 ncu --set full --clock-control none -k regex:"decode_kernel|scatter_kernel|apply_kernel|hidden_backward" -s 104 -c 4 -o /tmp/ncu/train python tools/profile_run.py --config B > /tmp/ncu/log_train.txt 2>&1
python tools/ncu_summary.py /tmp/ncu/train.ncu-rep > gpurun_out/c13_train_ncu_full.txt
echo "ncu train rc $?"
timeout 600 ncu --set full --clock-control none -k regex:"decode_kernel" -s 32 -c 1 -o /tmp/ncu/decD python tools/profile_run.py --config D > /tmp/ncu/log_D.txt 2>&1
python tools/ncu_summary.py /tmp/ncu/decD.ncu-rep > gpurun_out/c13_decode_D_ncu_full.txt
echo "ncu D rc $?"
ls -la gpurun_out
