"""A/B measurements of the last GPU call of round 2 (one process per library variant; CDAE_B200_LIB picks it).
  python tools/final_ab.py gen            synthetic sets of the three shapes -> /tmp/cdae_ab_*.npz (CPU only)
  python tools/final_ab.py fd [C] [E]     full-item-decode training, per-kernel-class device time and TFLOP/s
  python tools/final_ab.py topn           cdae_topn_build at config B's shape after 3 epochs, probe pass off / on
Every line printed is one JSON object (appended by the caller to gpurun_out/*.jsonl)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdae_b200 import CDAE, CDAEConfig, synth  # noqa: E402

SHAPES = {   # users = two minibatches of 148 x 128 users for the full-decode shapes
    "C": dict(U=37888, I=27000, K=200, mean=145.0),
    "E": dict(U=18944, I=100000, K=256, mean=50.0),
    "B": dict(U=100000, I=50000, K=50, mean=30.0),
}


def path(name):
    return "/tmp/cdae_ab_%s.npz" % name


def dataset(name):
    s = SHAPES[name]
    if os.path.exists(path(name)):
        d = np.load(path(name))
        return {k: d[k] for k in d.files}
    d = synth.make_dataset(s["U"], s["I"], mean_train=s["mean"], seed=1)
    return d


def gen():
    for name in sys.argv[2:] or ["C", "E", "B"]:
        s = SHAPES[name]
        t = time.time()
        d = synth.make_dataset(s["U"], s["I"], mean_train=s["mean"], seed=1)
        np.savez(path(name) + ".tmp.npz", train_row_ptr=d["train_row_ptr"], train_col=d["train_col"])
        os.replace(path(name) + ".tmp.npz", path(name))
        print(json.dumps(dict(gen=name, s=round(time.time() - t, 1), nnz=int(d["train_row_ptr"][-1]))), flush=True)


def fd(name):
    s = SHAPES[name]
    U, I, K = s["U"], s["I"], s["K"]
    d = dataset(name)
    cfg = CDAEConfig(loss="CE", num_dim=K, beta=1.0, asymmetric=True, corruption_ratio=0.5, scaled=True,
                     full_decode=True, batch_users=0)
    m = CDAE(cfg).reset(U, I, d["train_row_ptr"], d["train_col"])
    m.init_params(3)
    m.train_one_iteration(seed=1, epoch=0)
    plain = [m.train_one_iteration(seed=1, epoch=ep).device_ms for ep in range(1, 4)]
    m.profile(True)
    n_ep = 3
    for ep in range(4, 4 + n_ep):
        m.train_one_iteration(seed=1, epoch=ep)
    prof = m.profile_get()
    m.profile(False)
    per = {k: round(v[0] / n_ep, 4) for k, v in prof.items() if v[1]}
    flops = 2.0 * U * I * K
    tens = [k for k in ("fd_score", "fd_hidden", "fd_itemgrad") if k in per]
    out = dict(what="fd", lib=os.environ.get("CDAE_B200_LIB", "default"), shape=name, U=U, I=I, K=K,
               epoch_ms=[round(x, 4) for x in plain], users_per_s=U / (min(plain) * 1e-3), per_class_ms=per,
               tflops=dict({k: flops / (per[k] * 1e-3) / 1e12 for k in tens},
                           all3=3 * flops / (sum(per[k] for k in tens) * 1e-3) / 1e12,
                           step_6IK=3 * flops / (min(plain) * 1e-3) / 1e12))
    print(json.dumps(out), flush=True)
    m.close() if hasattr(m, "close") else None


def topn():
    s = SHAPES["B"]
    U, I, K = s["U"], s["I"], s["K"]
    d = dataset("B")
    cfg = CDAEConfig(loss="CE", num_dim=K, beta=1.0, corruption_ratio=0.5, scaled=True, num_neg=5, batch_users=16384)
    m = CDAE(cfg).reset(U, I, d["train_row_ptr"], d["train_col"])
    m.init_params(3)
    for ep in range(3):
        m.train_one_iteration(seed=1, epoch=ep)
    lists = {}
    for probe in ("0", "1", "8", "16", "2"):
        os.environ["CDAE_B200_TOPN_PROBE"] = probe
        m.pre_recommend(10)
        m.profile(True)
        reps = 3
        for _ in range(reps):
            m.pre_recommend(10)
        prof = m.profile_get()
        m.profile(False)
        ids, _ = m.recommend_all(10)
        lists[probe] = ids
        pth, verified, redone = m.topn_stats()
        keys = ("gather", "activate", "topn", "topn_pack", "topn_rerank", "topn_exact")
        out = dict(what="topn", lib=os.environ.get("CDAE_B200_LIB", "default"), probe=probe, probe_items=m.topn_probe_items(),
                   path=pth, verified=verified, redone=redone,
                   candidate_kernel_ms=round(prof["topn"][0] / reps, 4),
                   per_class_ms={k: round(prof[k][0] / reps, 4) for k in keys if k in prof and prof[k][1]},
                   users_per_s=U / (sum(prof[k][0] for k in keys if k in prof) / reps / 1e3))
        print(json.dumps(out), flush=True)
    print(json.dumps(dict(what="topn_lists_equal", equal=bool(all(np.array_equal(lists["0"], v) for v in lists.values())))), flush=True)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "gen":
        gen()
    elif mode == "fd":
        for nm in sys.argv[2:] or ["C"]:
            fd(nm)
    elif mode == "topn":
        topn()
