"""Trajectory (statistical) parity, SURVEY §8c tier 4: ranking quality after N epochs as a function
of the frozen-minibatch size, against (a) batch_users = 1 on the GPU, which is the reference's
per-user online step, and (b) the CPU oracle's sequential epoch (reference semantics, fp64).
Prints one JSON line per configuration."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    from cdae_b200 import CDAE, CDAEConfig, synth
    from oracle import oracle as orc
    U, I, K, epochs = int(os.environ.get("BQ_U", 12000)), int(os.environ.get("BQ_I", 4000)), 50, int(os.environ.get("BQ_EPOCHS", 30))
    d = synth.make_dataset(U, I, 30.0, seed=11)
    rp, col, trp, tcol = d["train_row_ptr"], d["train_col"], d["test_row_ptr"], d["test_col"]
    cfg = orc.default_config(loss="CE", num_dim=K, beta=1.0)
    names = ["P@1", "P@5", "P@10", "R@1", "R@5", "R@10", "MAP@5", "MAP@10"]
    batches = [int(x) for x in os.environ.get("BQ_BATCHES", "1,64,512,2048,8192,%d" % U).split(",") if x]
    every = int(os.environ.get("BQ_EVAL_EVERY", 5))
    for B in batches:
        if B <= 0:
            continue
        m = CDAE(CDAEConfig(batch_users=B, **cfg)).reset(U, I, rp, col)
        m.init_params(3)
        t = time.perf_counter()
        curve = []
        for e in range(epochs):
            st = m.train_one_iteration(seed=5, epoch=e)
            if e % every == every - 1 or e == epochs - 1:
                m.pre_recommend(10)
                met, n = m.topn_evaluate(trp, tcol)
                curve.append((e + 1, round(float(met[7]), 5), round(float(met[5]), 5)))
        out = {"impl": "gpu", "batch_users": B, "epochs": epochs, "train_s": round(time.perf_counter() - t, 2),
               "users": U, "items": I, "final": {k: round(float(v), 5) for k, v in zip(names, met)}, "map10_r10_curve": curve,
               "loss_last_epoch": st.loss_sum}
        print(json.dumps(out), flush=True)
        m.close()
    if os.environ.get("BQ_ORACLE", "1") == "1":
        o = orc.Oracle(cfg, U, I, rp, col)
        o.init_params(3)
        t = time.perf_counter()
        curve = []
        for e in range(epochs):
            o.train_epoch(5, e, batch_users=1)
            if e % every == every - 1 or e == epochs - 1:
                met, n = o.topn_evaluate(trp, tcol)
                curve.append((e + 1, round(float(met[7]), 5), round(float(met[5]), 5)))
                print(json.dumps({"impl": "oracle sequential", "epoch": e + 1, "map10": curve[-1][1], "r10": curve[-1][2]}), file=sys.stderr, flush=True)
        print(json.dumps({"impl": "oracle sequential (reference semantics, fp64, CPU)", "epochs": epochs, "users": U, "items": I,
                          "train_s": round(time.perf_counter() - t, 2),
                          "final": {k: round(float(v), 5) for k, v in zip(names, met)}, "map10_r10_curve": curve}), flush=True)


if __name__ == "__main__":
    main()
