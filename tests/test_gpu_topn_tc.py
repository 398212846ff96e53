"""The tcgen05/TMEM full-item decode (csrc/topn_tc.cuh) behind cdae_topn_build: lists must be
IDENTICAL to the oracle's CDAE::recommend (cdae.hpp:162-196) — the bf16 contraction only proposes
candidates, an error bound proves each user's list or sends the user to the exact kernel."""
import os

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orc(oracle_built):
    return oracle_built


def gpu_model(cfg, U, I, rp, col, params):
    from cdae_b200 import CDAE, CDAEConfig
    m = CDAE(CDAEConfig(**cfg)).reset(U, I, rp, col)
    m.set_params(params)
    return m


def check_lists(orc, cfg, data, params, topk=10, users=None):
    U, I = data["U"], data["I"]
    o = orc.Oracle(cfg, U, I, data["train_row_ptr"], data["train_col"])
    o.set_params(params)
    m = gpu_model(cfg, U, I, data["train_row_ptr"], data["train_col"], params)
    ids, sc = m.recommend_all(topk)
    stats = m.topn_stats()
    for u in (range(U) if users is None else users):
        oi, os_ = o.recommend(u, topk)
        assert ids[u].tolist() == oi.tolist(), (u, stats)
        np.testing.assert_allclose(sc[u], os_, rtol=1e-4, atol=1e-5)
    return m, stats


@pytest.mark.parametrize("K", [1, 10, 50, 62, 63, 100, 126, 200, 256, 318])
def test_every_k_block_count_matches_oracle(orc, K):
    """K + 2 bias columns fill 1..5 swizzle blocks of 64 (62/63 and 126 straddle a block edge)."""
    cfg = orc.default_config(loss="CE", num_dim=K, asymmetric=(K % 2 == 0))
    data = cases.small_dataset(U=300, I=1100, mean=14.0, seed=100 + K)
    p = cases.random_params(data["U"], data["I"], K, K, cfg["asymmetric"], True)
    m, (path, verified, redone) = check_lists(orc, cfg, data, p)
    assert path == 1 and verified + redone == data["U"]
    # the bound is not vacuous (the widest operands leave the smallest candidate buffers, 36 slots)
    assert verified >= (0.9 if K < 255 else 0.6) * data["U"]
    met, n = m.topn_evaluate(data["test_row_ptr"], data["test_col"])
    o = orc.Oracle(cfg, data["U"], data["I"], data["train_row_ptr"], data["train_col"])
    o.set_params(p)
    want, n2 = o.topn_evaluate(data["test_row_ptr"], data["test_col"])
    assert n == n2
    np.testing.assert_allclose(met, want, atol=1e-12)


def test_ragged_shapes(orc):
    """I below one tile, I one past a tile edge, U one past a CTA; heavy users (n_u ~ I/4)."""
    for U, I, mean, seed in [(37, 90, 9.0, 1), (129, 257, 12.0, 2), (257, 513, 40.0, 3), (5, 2000, 300.0, 4)]:
        cfg = orc.default_config(loss="SQUARE", num_dim=20)
        data = cases.small_dataset(U=U, I=I, mean=mean, seed=seed)
        p = cases.random_params(data["U"], data["I"], 20, seed, False, True)
        _, (path, verified, redone) = check_lists(orc, cfg, data, p)
        assert path == 1 and verified + redone == data["U"]


def test_trained_model_lists(orc):
    """After training, rated items dominate the top scores: the bitmap must keep them out."""
    cfg = orc.default_config(loss="CE", num_dim=50, beta=1.0)
    data = cases.small_dataset(U=400, I=900, mean=14.0, seed=9)
    U, I = data["U"], data["I"]
    p = cases.random_params(U, I, 50, 3, False, True, warm=False)
    from cdae_b200 import CDAE, CDAEConfig
    m = CDAE(CDAEConfig(batch_users=64, **cfg)).reset(U, I, data["train_row_ptr"], data["train_col"])
    m.set_params(p)
    for e in range(8):
        m.train_one_iteration(seed=4, epoch=e)
    trained = {k: v for k, v in m.get_params().items() if v.size}
    _, (path, verified, redone) = check_lists(orc, cfg, data, trained)
    assert path == 1 and verified + redone == U


def test_all_scores_tied_goes_to_the_exact_path(orc):
    """All-zero weights: the bound cannot separate ties, every user is redone exactly and the
    reference's tie rule (lowest unrated ids, heap.hpp:44-52) holds."""
    cfg = orc.default_config(loss="CE")
    data = cases.small_dataset(U=140, I=600, mean=9.0, seed=71)
    m = gpu_model(cfg, data["U"], data["I"], data["train_row_ptr"], data["train_col"], {})
    ids, _ = m.recommend_all(10)
    path, verified, redone = m.topn_stats()
    assert path == 1 and redone == data["U"] and verified == 0
    rp, col = data["train_row_ptr"], data["train_col"]
    for u in range(data["U"]):
        rated = set(col[rp[u]:rp[u + 1]].tolist())
        assert ids[u].tolist() == [i for i in range(data["I"]) if i not in rated][:10]


def test_near_ties_are_detected_not_guessed(orc):
    """Items that differ by less than bf16 resolution: identical rows except a 1e-6 bias ramp.  The
    approximate scores cannot order them; the lists must still be exact."""
    cfg = orc.default_config(loss="CE", num_dim=16)
    data = cases.small_dataset(U=64, I=400, mean=9.0, seed=5)
    U, I = data["U"], data["I"]
    rng = np.random.default_rng(0)
    w = rng.uniform(-0.3, 0.3, 16).astype(np.float32).astype(np.float64)
    p = {"W": np.tile(w, (I, 1)), "Wu": rng.uniform(-0.3, 0.3, (U, 16)).astype(np.float32).astype(np.float64),
         "b_prime": (np.arange(I)[::-1] * 1e-6).astype(np.float32).astype(np.float64)}
    check_lists(orc, cfg, data, p)


def test_fp32_path_selection(orc, monkeypatch):
    cfg = orc.default_config(loss="CE", num_dim=20)
    data = cases.small_dataset(U=100, I=400, mean=9.0, seed=8)
    p = cases.random_params(data["U"], data["I"], 20, 1, False, True)
    monkeypatch.setenv("CDAE_B200_TOPN", "fp32")
    _, (path, verified, redone) = check_lists(orc, cfg, data, p)
    assert path == 0
    monkeypatch.delenv("CDAE_B200_TOPN")
    _, (path, _, _) = check_lists(orc, cfg, data, p, topk=20)     # topk > 16: exact path
    assert path == 0
    cfg2 = orc.default_config(loss="CE", num_dim=320)              # K + 2 > 320: exact path
    p2 = cases.random_params(data["U"], data["I"], 320, 1, False, True)
    _, (path, _, _) = check_lists(orc, cfg2, data, p2, users=range(0, 100, 7))
    assert path == 0


def test_short_candidate_error_still_raised(orc):
    from cdae_b200 import CdaeError
    cfg = orc.default_config(loss="CE")
    rp2 = np.array([0, 8], np.int64)
    col2 = np.arange(8, dtype=np.int32)
    m2 = gpu_model(cfg, 1, 12, rp2, col2, {})
    with pytest.raises(CdaeError):
        m2.pre_recommend(10)
