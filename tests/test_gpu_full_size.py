"""BASELINE.json's configurations at (per-GPU) full size: size-independent properties of a training
epoch plus exact oracle checks on a sample of users — the oracle cannot run whole epochs at these
sizes in seconds, but it can score individual users against the trained parameters.

  B  100,000 x 50,000, K=50, 5 negatives, tied          (the benchmark configuration)
  C  27,000 items, K=200, ~145 items per user, asymmetric (user count cut to 20,000: the item
     side, the row lengths and the K=200 / 4-block tensor path are the full-size ones)
  D  1M x 200K, K=100 over 8 GPUs -> one rank's 125,000 users against all 200,000 items
  E  500K x 100K, K=256 over 8 GPUs -> one rank's 62,500 users against all 100,000 items
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CONFIGS = {
    "B": dict(U=100_000, I=50_000, mean=30.0, K=50, asym=False),
    "C": dict(U=20_000, I=27_000, mean=145.0, K=200, asym=True),
    "D": dict(U=125_000, I=200_000, mean=30.0, K=100, asym=False),
    "E": dict(U=62_500, I=100_000, mean=50.0, K=256, asym=True),
}


@pytest.mark.parametrize("name", ["B", "C", "D", "E"])
def test_full_size_epoch_and_recommend(oracle_built, name):
    from cdae_b200 import CDAE, CDAEConfig, synth
    orc = oracle_built
    c = CONFIGS[name]
    U, I, K = c["U"], c["I"], c["K"]
    d = synth.make_dataset(U, I, c["mean"], seed=20141119)
    rp, col = d["train_row_ptr"], d["train_col"]
    nnz = len(col)
    cfg = orc.default_config(loss="CE", num_dim=K, beta=1.0, asymmetric=c["asym"])
    m = CDAE(CDAEConfig(**cfg)).reset(U, I, rp, col)
    m.init_params(5)

    # ---- one pass = every user once, every positive with its 5 negatives, ~half the inputs kept
    losses = []
    for epoch in range(2):
        st = m.train_one_iteration(seed=9, epoch=epoch)
        assert st.user_steps == U
        assert st.outputs == nnz * (1 + cfg["num_neg"])
        assert abs(st.inputs_kept - 0.5 * nnz) < 6 * np.sqrt(0.25 * nnz)     # Binomial(nnz, 1/2)
        assert np.isfinite(st.loss_sum)
        losses.append(st.loss_sum)
    assert losses[1] < losses[0]                                               # it learns

    # ---- recommend on the tensor path; every list accounted for
    ids, sc = m.recommend_all(10)
    path, verified, redone = m.topn_stats()
    assert path == 1 and verified + redone == U
    assert (ids >= 0).all() and (ids < I).all()
    assert (np.diff(sc, axis=1) <= 0).all()                                    # sorted by score
    # ---- exact checks on a sample of users against the oracle with the trained parameters
    params = {k: v for k, v in m.get_params().items() if v.size}
    o = orc.Oracle(cfg, U, I, rp, col)
    o.set_params(params)
    rng = np.random.default_rng(1)
    heavy = int(np.argmax(np.diff(rp)))                                        # longest row too
    sample = np.unique(np.concatenate([rng.choice(U, 24, replace=False), [0, U - 1, heavy]]))
    z = m.encode(sample)
    for j, u in enumerate(sample):
        row = col[rp[u]:rp[u + 1]]
        assert not np.isin(ids[u], row).any()                                  # rated items excluded
        want_z = o.hidden(int(u), row.astype(np.int64), 1.0)
        np.testing.assert_allclose(z[j], want_z, rtol=1e-4, atol=1e-6)         # north_star: 1e-4 on hidden
        oi, os_ = o.recommend(int(u), 10)
        assert ids[u].tolist() == oi.tolist(), (name, u)                       # identical top-N
        np.testing.assert_allclose(sc[u], os_, rtol=1e-4, atol=1e-5)
    # ---- idempotence: rebuilding the lists from the same parameters gives the same table
    m.pre_recommend(10)
    ids2, _ = m.recommend_all(10)
    assert np.array_equal(ids, ids2)
    m.close()
