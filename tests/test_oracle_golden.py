"""The C restatement against the committed golden vectors (generated from the verbatim
reference by tests/golden/make_golden.py).  CPU only; runs on the GPU box too, where
/root/reference does not exist."""
import numpy as np
import pytest

from tests import golden


@pytest.fixture(scope="module")
def orc(oracle_built):
    return oracle_built


def make_oracle(orc, g):
    o = orc.Oracle(g["cfg"], g["U"], g["I"], g["train_row_ptr"], g["train_col"])
    o.set_params(g["p0"])
    return o


def test_golden_files_present():
    assert len(golden.NAMES) >= 5


@pytest.mark.parametrize("name", golden.NAMES)
def test_forward_and_topn(orc, name):
    g = golden.load(name)
    o = make_oracle(orc, g)
    cfg, rp, col = g["cfg"], g["train_row_ptr"], g["train_col"]
    scale = 1.0 / (1.0 - cfg["corruption_ratio"]) if cfg["scaled"] else 1.0
    for u in range(g["U"]):
        row = col[rp[u]:rp[u + 1]]
        keep = g["keep"][rp[u]:rp[u + 1]].astype(bool)
        np.testing.assert_allclose(o.hidden(u, row), g["z_clean"][u], rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(o.hidden(u, row[keep], scale), g["z_corrupt"][u], rtol=1e-12,
                                   atol=1e-300)
        assert o.recommend(u, 10)[0].tolist() == g["rec_before"][u].tolist()
    assert abs(o.penalty_loss() - g["penalty_before"]) <= 1e-12 * abs(g["penalty_before"])
    if not np.isnan(g["data_loss_q0"]):
        assert abs(o.data_loss(None) - g["data_loss_q0"]) <= 1e-10 * abs(g["data_loss_q0"])
    m, _ = o.topn_evaluate(g["test_row_ptr"], g["test_col"])
    np.testing.assert_allclose(m, g["metrics_before"], atol=6e-5)


@pytest.mark.parametrize("name", golden.NAMES)
def test_sequential_pass(orc, name):
    g = golden.load(name)
    o = make_oracle(orc, g)
    cfg, rp, col = g["cfg"], g["train_row_ptr"], g["train_col"]
    nu = cfg["num_neg"]
    for u in range(g["U"]):
        row = col[rp[u]:rp[u + 1]]
        keep = g["keep"][rp[u]:rp[u + 1]].astype(bool)
        o.step_sequential(u, row[keep], g["negs"][rp[u] * nu:rp[u + 1] * nu])
    for k, want in g["p1"].items():
        np.testing.assert_allclose(o.param(k), want, rtol=1e-9, atol=1e-12, err_msg=k)
    for u in range(g["U"]):
        assert o.recommend(u, 10)[0].tolist() == g["rec_after"][u].tolist()
    m, _ = o.topn_evaluate(g["test_row_ptr"], g["test_col"])
    np.testing.assert_allclose(m, g["metrics_after"], atol=6e-5)


@pytest.mark.parametrize("name", golden.NAMES)
def test_frozen_batch_of_one_reproduces_reference(orc, name):
    """Negatives in the golden files are unique per user, so |B|=1 frozen == the online step."""
    g = golden.load(name)
    o = make_oracle(orc, g)
    cfg, rp, col = g["cfg"], g["train_row_ptr"], g["train_col"]
    nu = cfg["num_neg"]
    for u in range(g["U"]):
        row = col[rp[u]:rp[u + 1]]
        keep = g["keep"][rp[u]:rp[u + 1]].astype(bool)
        o.step_frozen([u], [row[keep]], [g["negs"][rp[u] * nu:rp[u + 1] * nu]])
    for k, want in g["p1"].items():
        np.testing.assert_allclose(o.param(k), want, rtol=1e-9, atol=1e-12, err_msg=k)
