"""First-contact check of the tcgen05 top-N kernel against the fp32 path (single GPU)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def run(U, I, K, mean, seed=3):
    from cdae_b200 import CDAE, CDAEConfig, synth
    d = synth.make_dataset(U, I, mean, seed=seed)
    rp, col = d["train_row_ptr"], d["train_col"]
    out = {}
    for path in ("fp32", "tc"):
        os.environ["CDAE_B200_TOPN"] = path
        m = CDAE(CDAEConfig(loss="CE", num_dim=K, beta=1.0)).reset(U, I, rp, col)
        m.init_params(7)
        rng = np.random.default_rng(1)
        m.set_params({"b_prime": rng.uniform(-0.05, 0.05, I), "b": rng.uniform(-0.2, 0.2, K)})
        t = time.perf_counter()
        m.pre_recommend(10)
        dt = time.perf_counter() - t
        m.profile(True)
        m.pre_recommend(10)
        prof = m.profile_get()
        ids, sc = m.recommend_all(10)
        out[path] = (ids, sc, m.topn_stats(), dt, prof)
        m.close()
    same = (out["fp32"][0] == out["tc"][0]).all(axis=1)
    tc_ms = out["tc"][4]["topn"][0] or 1e-9
    print("   tensor path: %.1f TFLOP/s algorithmic (2*U*I*K), %.1f executed (padded K)"
          % (2.0 * U * I * K / tc_ms / 1e9, 2.0 * U * I * ((K + 2 + 63) // 64 * 64) / tc_ms / 1e9), flush=True)
    print("U=%d I=%d K=%d: identical lists %d/%d  tc stats(path,verified,redone)=%s  first-call wall fp32 %.3fs tc %.3fs"
          % (U, I, K, same.sum(), U, out["tc"][2], out["fp32"][3], out["tc"][3]), flush=True)
    for p in ("fp32", "tc"):
        print("   %s kernel ms: %s" % (p, {k: round(v[0], 3) for k, v in out[p][4].items() if v[1]}), flush=True)
    if not same.all():
        bad = np.where(~same)[0][:3]
        for u in bad:
            print("   user", u, out["fp32"][0][u], out["tc"][0][u], flush=True)
    return bool(same.all())


if __name__ == "__main__":
    ok = True
    for args in ((300, 1000, 50, 12.0), (1000, 5000, 100, 20.0), (2000, 3000, 200, 20.0), (700, 2500, 256, 20.0)):
        ok &= run(*args)
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        ok &= run(100000, 50000, 50, 30.0)
    if len(sys.argv) > 1 and sys.argv[1] == "shapes":      # BASELINE.json configs C, D (one rank), E (one rank)
        ok &= run(20000, 27000, 200, 145.0)
        ok &= run(125000, 200000, 100, 30.0)
        ok &= run(62500, 100000, 256, 50.0)
    print("tc_probe:", "ok" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)
