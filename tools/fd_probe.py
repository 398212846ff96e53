"""Full-item-decode training at a BASELINE shape: per-kernel-class device time and TFLOP/s.
  python tools/fd_probe.py [U I K mean [batch_users]]     (default: 2 minibatches of config C's shape)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdae_b200 import CDAE, CDAEConfig, synth  # noqa: E402


def main():
    a = sys.argv[1:]
    U, I, K = (int(a[0]), int(a[1]), int(a[2])) if len(a) >= 3 else (37888, 27000, 200)
    mean = float(a[3]) if len(a) >= 4 else 145.0
    bu = int(a[4]) if len(a) >= 5 else 0
    t = time.time()
    d = synth.make_dataset(U, I, mean_train=mean, seed=1)
    print("dataset %.1fs nnz %d" % (time.time() - t, d["train_row_ptr"][-1]), flush=True)
    cfg = CDAEConfig(loss="CE", num_dim=K, beta=1.0, asymmetric=True, corruption_ratio=0.5, scaled=True,
                     full_decode=True, batch_users=bu)
    m = CDAE(cfg).reset(U, I, d["train_row_ptr"], d["train_col"])
    m.init_params(3)
    m.train_one_iteration(seed=1, epoch=0)
    m.profile(True)
    ms = []
    for ep in range(1, 4):
        st = m.train_one_iteration(seed=1, epoch=ep)
        ms.append(st.device_ms)
    prof = m.profile_get()
    m.profile(False)
    n_ep = 3
    per = {k: (v[0] / n_ep, v[1] // n_ep) for k, v in prof.items() if v[1]}
    flops = 2.0 * U * I * K
    tens = [k for k in ("fd_score", "fd_hidden", "fd_itemgrad") if k in per]
    share = {"fd_score": 2.0 if "fd_hidden" not in per else 1.0, "fd_hidden": 1.0, "fd_itemgrad": 1.0}   # fused: score does 2 of the 3 contractions
    out = dict(U=U, I=I, K=K, epoch_ms=ms, users_per_s=U / (min(ms) * 1e-3), per_class_ms=per,
               tflops=dict({k: share[k] * flops / (per[k][0] * 1e-3) / 1e12 for k in tens},
                           all3=3 * flops / (sum(per[k][0] for k in tens) * 1e-3) / 1e12,
                           step_6IK=3 * flops / (min(ms) * 1e-3) / 1e12))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
