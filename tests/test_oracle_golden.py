"""The C restatement against the committed golden vectors (generated from the verbatim
reference by tests/golden/make_golden.py).  CPU only; runs on the GPU box too, where
/root/reference does not exist."""
import numpy as np
import pytest

from tests import cases, golden


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


@pytest.mark.parametrize("kw", [dict(loss="CE", beta=1.0), dict(loss="SQUARE", asymmetric=True),
                                dict(loss="CE", asymmetric=True, linear_function=True, tanh=True),
                                dict(loss="CE", using_adagrad=False, user_factor=False)])
def test_full_decode_mode_is_frozen_step_with_all_negatives(oracle_built, kw):
    """H12 has no reference function: the oracle's full-item-decode step must be exactly its
    frozen-batch step (pinned to the verbatim reference above / in test_oracle_vs_reference.py)
    called with every non-positive item as a negative, once.  The bf16-rounding variant used to
    check the tensor-core kernels must stay within bf16 resolution of it."""
    orc = oracle_built
    cfg = orc.default_config(num_dim=12, **kw)
    data = cases.small_dataset(U=24, I=70, mean=6.0, seed=3)
    U, I, rp, col = data["U"], data["I"], data["train_row_ptr"], data["train_col"]
    p = cases.random_params(U, I, 12, 1, cfg["asymmetric"], cfg["user_factor"], cfg["linear_function"])
    a, b, c = (orc.Oracle(cfg, U, I, rp, col) for _ in range(3))
    for o in (a, b, c):
        o.set_params(p)
    users = list(range(U))
    ins = [col[rp[u]:rp[u + 1]][a.sample_keep(5, 0, u).astype(bool)] for u in users]
    negs = [np.setdiff1d(np.arange(I), col[rp[u]:rp[u + 1]]) for u in users]
    la = a.step_frozen(users, ins, negs)
    lb = b.step_frozen_full(users, ins, rounding=0)
    c.step_frozen_full(users, ins, rounding=1)
    assert abs(la - lb) <= 1e-12 * abs(la)
    for k, v in a.get_params().items():
        np.testing.assert_allclose(b.param(k), v, rtol=1e-12, atol=1e-13, err_msg=k)
        if not k.endswith("_ag"):
            np.testing.assert_allclose(c.param(k), v, rtol=3e-2, atol=1e-2, err_msg=k)
    # epoch driver == explicit minibatches with the Philox masks
    d, e = (orc.Oracle(cfg, U, I, rp, col) for _ in range(2))
    for o in (d, e):
        o.set_params(p)
    d.train_epoch_full(9, 1, 10, rounding=0)
    for b0 in range(0, U, 10):
        us = list(range(b0, min(U, b0 + 10)))
        e.step_frozen_full(us, [col[rp[u]:rp[u + 1]][e.sample_keep(9, 1, u).astype(bool)] for u in us], rounding=0)
    for k, v in d.get_params().items():
        np.testing.assert_array_equal(e.param(k), v, err_msg=k)
