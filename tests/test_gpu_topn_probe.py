"""Probe pass of cdae_topn_build (csrc/topn_tc.cuh "Probe pass", topn_api.inl tc_probe_thresholds): the first
sweep of the tensor-core candidate kernel starts from per-user thresholds derived from the items with the
largest mean-user score instead of -inf.  Only the cost of the build may depend on it: the lists must stay
IDENTICAL to the oracle's CDAE::recommend (cdae.hpp:162-196), and identical to the lists without the probe."""
import numpy as np
import pytest

from tests import cases
from tests.test_gpu_topn_tc import check_lists, gpu_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orc(oracle_built):
    return oracle_built


@pytest.fixture()
def probe_on(monkeypatch):
    monkeypatch.setenv("CDAE_B200_TOPN_PROBE", "1")


def probe_size(I, knob=1):
    """tc_probe_thresholds (topn_api.inl): 256-item tiles, one per 2048 items, at most 4 (knob n >= 2: at most
    min(n, 32))."""
    if I < 512 or knob <= 0:
        return 0
    return 256 * min(4 if knob == 1 else min(knob, 32), max(1, I // 2048))


@pytest.mark.parametrize("K,I,U,mean", [(10, 1100, 300, 14.0), (50, 1100, 300, 14.0), (100, 2600, 300, 14.0),
                                        (256, 1100, 300, 14.0), (50, 5000, 600, 30.0), (20, 9000, 900, 40.0)])
def test_lists_match_oracle_with_probe(orc, probe_on, K, I, U, mean):
    cfg = orc.default_config(loss="CE", num_dim=K, asymmetric=(K % 20 == 0))
    data = cases.small_dataset(U=U, I=I, mean=mean, seed=300 + K)
    p = cases.random_params(data["U"], data["I"], K, K, cfg["asymmetric"], True)
    m, (path, verified, redone) = check_lists(orc, cfg, data, p, users=range(0, data["U"], max(1, data["U"] // 300)))
    assert path == 1 and verified + redone == data["U"]
    assert m.topn_probe_items() == probe_size(data["I"])
    assert verified >= (0.9 if K < 255 else 0.6) * data["U"]


def test_probe_is_skipped_for_small_item_tables(orc, probe_on):
    cfg = orc.default_config(loss="CE", num_dim=20)
    data = cases.small_dataset(U=100, I=400, mean=9.0, seed=8)
    p = cases.random_params(data["U"], data["I"], 20, 1, False, True)
    m, (path, _, _) = check_lists(orc, cfg, data, p)
    assert path == 1 and m.topn_probe_items() == 0


def test_probe_tile_knob(orc, monkeypatch):
    """CDAE_B200_TOPN_PROBE=n >= 2 allows up to n tiles; the lists do not depend on it."""
    cfg = orc.default_config(loss="CE", num_dim=20)
    data = cases.small_dataset(U=900, I=9000, mean=40.0, seed=77)
    p = cases.random_params(data["U"], data["I"], 20, 2, False, True)
    ref = None
    for knob in ("0", "1", "2", "8"):
        monkeypatch.setenv("CDAE_B200_TOPN_PROBE", knob)
        m = gpu_model(cfg, data["U"], data["I"], data["train_row_ptr"], data["train_col"], p)
        ids, _ = m.recommend_all(10)
        assert m.topn_probe_items() == (0 if knob == "0" else probe_size(data["I"], int(knob)))
        ref = ids if ref is None else ref
        assert np.array_equal(ids, ref), knob


def test_trained_model_same_lists_with_and_without_probe(orc, monkeypatch):
    """After training the popular items lead every list (and are what most users have rated): the probe table is
    exactly those items.  Lists with the probe == lists without == oracle."""
    cfg = orc.default_config(loss="CE", num_dim=50, beta=1.0)
    data = cases.small_dataset(U=400, I=1500, mean=14.0, seed=19)
    U, I = data["U"], data["I"]
    p = cases.random_params(U, I, 50, 3, False, True, warm=False)
    from cdae_b200 import CDAE, CDAEConfig
    m = CDAE(CDAEConfig(batch_users=64, **cfg)).reset(U, I, data["train_row_ptr"], data["train_col"])
    m.set_params(p)
    for e in range(8):
        m.train_one_iteration(seed=4, epoch=e)
    trained = {k: v for k, v in m.get_params().items() if v.size}
    monkeypatch.setenv("CDAE_B200_TOPN_PROBE", "0")
    m0 = gpu_model(cfg, U, I, data["train_row_ptr"], data["train_col"], trained)
    ids0, sc0 = m0.recommend_all(10)
    assert m0.topn_probe_items() == 0
    monkeypatch.setenv("CDAE_B200_TOPN_PROBE", "1")
    m1, (path, verified, redone) = check_lists(orc, cfg, data, trained)
    assert path == 1 and verified + redone == U and m1.topn_probe_items() == probe_size(I)
    ids1, sc1 = m1.recommend_all(10)
    assert np.array_equal(ids0, ids1)
    np.testing.assert_allclose(sc0, sc1, rtol=0, atol=0)


def test_all_scores_tied_with_probe(orc, probe_on):
    """All-zero weights: every key and every score ties; the probe must not invent a threshold."""
    cfg = orc.default_config(loss="CE")
    data = cases.small_dataset(U=140, I=600, mean=9.0, seed=71)
    m = gpu_model(cfg, data["U"], data["I"], data["train_row_ptr"], data["train_col"], {})
    ids, _ = m.recommend_all(10)
    path, verified, redone = m.topn_stats()
    assert path == 1 and verified + redone == data["U"]
    rp, col = data["train_row_ptr"], data["train_col"]
    for u in range(data["U"]):
        rated = set(col[rp[u]:rp[u + 1]].tolist())
        assert ids[u].tolist() == [i for i in range(data["I"]) if i not in rated][:10]


def test_heavy_users_rated_most_of_the_probe_table(orc, probe_on):
    """Users whose train rows cover most of the probe items: fewer than k unrated probe items -> no start
    threshold for them (thr0 = -inf), lists still exact."""
    cfg = orc.default_config(loss="SQUARE", num_dim=20)
    data = cases.small_dataset(U=24, I=640, mean=150.0, seed=4)
    p = cases.random_params(data["U"], data["I"], 20, 4, False, True)
    _, (path, verified, redone) = check_lists(orc, cfg, data, p)
    assert path == 1 and verified + redone == data["U"]
