"""A short run of one BASELINE configuration for `ncu` (never a bench number): two warm epochs, then
`--epochs` epochs (and optionally one recommend pass).  ncu selects kernels with -k / -s / -c.
usage (on the GPU box):
  ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 26 -c 2 -o gpurun_out/x \
      python tools/profile_run.py --config B"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="B")
    ap.add_argument("--users", type=int, default=0, help="cut the user count (0 = the configuration's)")
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--warm", type=int, default=2)
    ap.add_argument("--topn", action="store_true")
    a = ap.parse_args()
    from cdae_b200 import CDAE, CDAEConfig, synth
    c = bench.CONFIGS[a.config]
    U = a.users or c["users_per_gpu"]
    d = synth.make_blocked_dataset(U, c["items"], c["mean"], seed=bench.SEED)
    m = CDAE(CDAEConfig(full_decode=c["full"], batch_users=c["batch"], **bench.model_cfg(c))).reset(
        U, c["items"], d["train_row_ptr"], d["train_col"])
    m.init_params(bench.SEED)
    for ep in range(a.warm + a.epochs):
        st = m.train_one_iteration(seed=bench.SEED, epoch=ep)
    print("epoch device ms", st.device_ms, "launches", st.kernel_launches)
    if a.topn:
        m.pre_recommend(10)
        m.pre_recommend(10)
    m.close()


if __name__ == "__main__":
    main()
