"""Generates tests/golden/*.npz from the VERBATIM reference (oracle/_ref/libcdae_ref.so).

Run in the dev container (needs /root/reference):   python tests/golden/make_golden.py
The reference cannot travel to the GPU box, its outputs can: every file holds the inputs
(CSR, config, parameters, explicit corruption masks and negatives) and what the reference
computed from them (hidden vectors, parameters + AdaGrad state after a full sequential pass of
train_one_user_corruption, top-10 lists, losses, top-N metrics).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402
from tests import cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

GOLDEN_CASES = {
    # name: (dataset kwargs, config overrides)
    "tied_ce_adagrad_k10": (dict(U=48, I=160, mean=10.0, seed=11), dict(loss="CE", num_neg=2)),
    "asym_square_sgd_k50": (dict(U=40, I=200, mean=12.0, seed=12),
                            dict(loss="SQUARE", asymmetric=True, using_adagrad=False, num_dim=50,
                                 num_neg=3, corruption_ratio=0.3)),
    "yelp_defaults_k50": (dict(U=40, I=900, mean=10.0, seed=13),   # apps/yelp flag defaults + cdae.sh
                          dict(loss="SQUARE", num_dim=50, num_neg=5, corruption_ratio=0.0,
                               scaled=False, beta=1.0)),
    "tanh_linfn_ce_k16": (dict(U=32, I=150, mean=9.0, seed=14),
                          dict(loss="CE", tanh=True, linear_function=True, beta=1.0, num_dim=16,
                               num_neg=1, corruption_ratio=0.6)),
    "nouser_linear_k8": (dict(U=32, I=140, mean=9.0, seed=15),
                         dict(loss="SQUARE", user_factor=False, linear=True, num_dim=8, num_neg=2,
                              corruption_ratio=0.5, scaled=True)),
}


def make_case(name):
    dkw, ckw = GOLDEN_CASES[name]
    data = cases.small_dataset(**dkw)
    cfg = orc.default_config(**ckw)
    U, I, K = data["U"], data["I"], cfg["num_dim"]
    rng = np.random.default_rng(1000 + dkw["seed"])
    params = cases.random_params(U, I, K, 2000 + dkw["seed"], cfg["asymmetric"], cfg["user_factor"],
                                 cfg["linear_function"])
    steps = cases.draw_step_inputs(data, cfg["num_neg"], cfg["corruption_ratio"], rng,
                                   unique_negs=True)
    rp, col = data["train_row_ptr"], data["train_col"]
    keep = np.concatenate([steps[u][0] for u in range(U)]).astype(np.uint8)
    negs = np.concatenate([steps[u][1] for u in range(U)]).astype(np.int32)

    ref = orc.Reference(cfg, U, I, rp, col)
    ref.set_params(params)
    scale = 1.0 / (1.0 - cfg["corruption_ratio"]) if cfg["scaled"] else 1.0
    z_clean = np.stack([ref.hidden(u, col[rp[u]:rp[u + 1]]) for u in range(U)])
    z_corrupt = np.stack([ref.hidden(u, col[rp[u]:rp[u + 1]][steps[u][0]], scale) for u in range(U)])
    rec_before = np.stack([ref.recommend(u, 10) for u in range(U)])
    penalty_before = ref.penalty_loss()
    # full-keep data loss is only deterministic when nothing is corrupted (q = 0)
    data_loss_q0 = ref.data_loss() if cfg["corruption_ratio"] == 0.0 else np.nan
    metrics_before = ref.topn_evaluate(data["test_row_ptr"], data["test_col"])
    for u in range(U):
        k, n = steps[u]
        ref.train_one_user(u, col[rp[u]:rp[u + 1]][k], n)
    after = ref.get_params()
    rec_after = np.stack([ref.recommend(u, 10) for u in range(U)])
    metrics_after = ref.topn_evaluate(data["test_row_ptr"], data["test_col"])

    out = dict(cfg=json.dumps(cfg), U=U, I=I, train_row_ptr=rp, train_col=col,
               test_row_ptr=data["test_row_ptr"], test_col=data["test_col"], keep=keep, negs=negs,
               z_clean=z_clean, z_corrupt=z_corrupt, rec_before=rec_before, rec_after=rec_after,
               penalty_before=penalty_before, data_loss_q0=data_loss_q0,
               metrics_before=metrics_before, metrics_after=metrics_after)
    for k, v in params.items():
        out["p0_" + k] = v
    for k, v in after.items():
        if v.size:
            out["p1_" + k] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return out


if __name__ == "__main__":
    orc.build(ref=True)
    assert orc.have_reference(), "needs /root/reference to build oracle/_ref"
    for name in GOLDEN_CASES:
        o = make_case(name)
        print(name, "U=%d I=%d nnz=%d" % (o["U"], o["I"], len(o["train_col"])))
