"""Trajectory (statistical) parity of full-item-decode training (H12), SURVEY §8c tier 4: ranking
quality after N epochs of the tcgen05 path (bf16 operands) against the CPU oracle's full-decode
epochs in plain fp64 (rounding = 0) and with the restated bf16 rounding (rounding = 1), same frozen
minibatches and Philox masks.  Also prints the sampled-negative path for context.
One JSON line per run."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    from cdae_b200 import CDAE, CDAEConfig, synth
    from oracle import oracle as orc
    U, I, K = int(os.environ.get("FQ_U", 3000)), int(os.environ.get("FQ_I", 2000)), int(os.environ.get("FQ_K", 50))
    epochs, B = int(os.environ.get("FQ_EPOCHS", 10)), int(os.environ.get("FQ_B", 512))
    d = synth.make_dataset(U, I, 30.0, seed=11)
    rp, col, trp, tcol = d["train_row_ptr"], d["train_col"], d["test_row_ptr"], d["test_col"]
    cfg = orc.default_config(loss="CE", num_dim=K, beta=1.0, asymmetric=True, learn_rate=0.05)
    names = ["P@1", "P@5", "P@10", "R@1", "R@5", "R@10", "MAP@5", "MAP@10"]

    def report(impl, met, t):
        print(json.dumps({"impl": impl, "users": U, "items": I, "K": K, "epochs": epochs, "batch_users": B,
                          "train_s": round(t, 2), "final": {k: round(float(v), 5) for k, v in zip(names, met)}}), flush=True)

    for full in (True, False):
        m = CDAE(CDAEConfig(batch_users=B, full_decode=full, **cfg)).reset(U, I, rp, col)
        m.init_params(3)
        t = time.perf_counter()
        for e in range(epochs):
            m.train_one_iteration(seed=5, epoch=e)
        t = time.perf_counter() - t
        m.pre_recommend(10)
        met, _ = m.topn_evaluate(trp, tcol)
        report("gpu full-item decode (bf16 tcgen05)" if full else "gpu sampled negatives (num_neg=5, fp32)", met, t)
        m.close()
    for rounding in (0, 1):
        o = orc.Oracle(cfg, U, I, rp, col)
        o.init_params(3)
        t = time.perf_counter()
        for e in range(epochs):
            o.train_epoch_full(5, e, B, rounding=rounding)
        t = time.perf_counter() - t
        met, _ = o.topn_evaluate(trp, tcol)
        report("oracle full-item decode, CPU fp64" + (" with bf16 operand rounding" if rounding else ""), met, t)


if __name__ == "__main__":
    main()
