"""Pins the C restatement (oracle/cdae_oracle.c) against the VERBATIM reference headers
compiled into oracle/_ref/libcdae_ref.so (oracle/ref_driver.cpp).

The reference's own test-suite has no CDAE test; this file is the
known-answer suite the oracle needs before anything is compared with it.  CPU only.
"""
import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.skipif(False, reason="")


@pytest.fixture(scope="module")
def orc(oracle_built):
    if not oracle_built.have_reference():
        pytest.skip("oracle/_ref/libcdae_ref.so not built (needs /root/reference)")
    return oracle_built


@pytest.fixture(scope="module")
def data():
    return cases.small_dataset()


def make_pair(orc, data, seed=3, **kw):
    kw.setdefault("loss", "CE")
    cfg = orc.default_config(**kw)
    U, I, K = data["U"], data["I"], cfg["num_dim"]
    ref = orc.Reference(cfg, U, I, data["train_row_ptr"], data["train_col"])
    o = orc.Oracle(cfg, U, I, data["train_row_ptr"], data["train_col"])
    p = cases.random_params(U, I, K, seed, cfg["asymmetric"], cfg["user_factor"],
                            cfg["linear_function"])
    ref.set_params(p)
    o.set_params(p)
    return cfg, ref, o


def assert_params_close(ref, o, tol=1e-12):
    for name in orc_names():
        a, b = ref.get_param(name), o.param(name)
        if a.size == 0:
            continue
        assert a.shape == b.shape, name
        err = np.max(np.abs(a - b) / (1e-300 + np.maximum(1.0, np.abs(a))))
        assert err <= tol, (name, err)


def orc_names():
    from oracle.oracle import PARAMS
    return PARAMS


@pytest.mark.parametrize("loss", ["SQUARE", "CE", "LOG", "HINGE", "SQUARED_HINGE", "LOGM"])
def test_loss_functions_bit_exact(orc, loss):
    preds = [-40., -18.000001, -18., -5., -1., -1e-9, 0., 1e-9, .3, 1., 1.000001, 7., 18., 18.5, 60.]
    for t in (0., 1., -1.):
        for y in preds:
            assert orc.loss_gradient(loss, y, t) == orc.ref_loss_gradient(loss, y, t)
            assert orc.loss_evaluate(loss, y, t) == orc.ref_loss_evaluate(loss, y, t)


def test_logistic_loss_inside_domain(orc):
    for y in (1e-6, .25, .5, .999):
        for t in (0., 1.):
            assert orc.loss_gradient("LOGISTIC", y, t) == orc.ref_loss_gradient("LOGISTIC", y, t)
            assert orc.loss_evaluate("LOGISTIC", y, t) == orc.ref_loss_evaluate("LOGISTIC", y, t)
    assert np.isnan(orc.loss_gradient("LOGISTIC", 1.5, 1.))   # the reference CHECK-aborts here


@pytest.mark.parametrize("kw", cases.CONFIG_GRID)
def test_hidden_and_output(orc, data, kw):
    cfg, ref, o = make_pair(orc, data, **kw)
    rng = np.random.default_rng(0)
    rp, col = data["train_row_ptr"], data["train_col"]
    for u in (0, 5, data["U"] - 1):
        row = col[rp[u]:rp[u + 1]].astype(np.int64)
        for items, scale in ((row, 1.0), (row[rng.random(len(row)) > .5], 2.0), (row[:0], 1.0)):
            # same summation order on both sides: the reference iterates a hash map, so feed
            # the oracle the items in the order that map yields them
            order = [i for i in ref.user_items_order(u) if i in set(items.tolist())]
            zr = ref.hidden(u, items, scale)
            zo = o.hidden(u, np.array(order, np.int64), scale)
            np.testing.assert_allclose(zo, zr, rtol=1e-14, atol=1e-300)
            for it in (0, 17, data["I"] - 1):
                assert abs(o.output(zr, it) - ref.output(zr, it)) <= 1e-14 * max(1, abs(ref.output(zr, it)))


@pytest.mark.parametrize("kw", cases.CONFIG_GRID)
def test_sequential_steps_match_reference(orc, data, kw):
    """train_one_user_corruption over every user, in uid order, explicit masks/negatives —
    parameters and AdaGrad state must agree after the whole pass."""
    cfg, ref, o = make_pair(orc, data, **kw)
    rng = np.random.default_rng(11)
    steps = cases.draw_step_inputs(data, cfg["num_neg"], cfg["corruption_ratio"], rng)
    rp, col = data["train_row_ptr"], data["train_col"]
    for u in range(data["U"]):
        keep, negs = steps[u]
        row = col[rp[u]:rp[u + 1]].astype(np.int64)
        in_items = row[keep]
        ref.train_one_user(u, in_items, negs)
        o.step_sequential(u, in_items, negs, out_order=ref.user_items_order(u))
    assert_params_close(ref, o, 1e-11)
    # and the CSR visiting order changes nothing beyond rounding
    cfg2, ref2, o2 = make_pair(orc, data, **kw)
    for u in range(data["U"]):
        keep, negs = steps[u]
        row = col[rp[u]:rp[u + 1]].astype(np.int64)
        o2.step_sequential(u, row[keep], negs)
    for name in orc_names():
        np.testing.assert_allclose(o2.param(name), o.param(name), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("kw", [dict(num_neg=2), dict(num_neg=2, asymmetric=True, loss="SQUARE"),
                                dict(num_neg=1, linear_function=True, beta=1.0)])
def test_frozen_batch_of_one_equals_sequential(orc, data, kw):
    """|B| = 1 frozen-batch == the reference step when no negative repeats."""
    cfg, ref, o = make_pair(orc, data, **kw)
    rng = np.random.default_rng(5)
    steps = cases.draw_step_inputs(data, cfg["num_neg"], cfg["corruption_ratio"], rng,
                                   unique_negs=True)
    rp, col = data["train_row_ptr"], data["train_col"]
    for u in range(data["U"]):
        keep, negs = steps[u]
        row = col[rp[u]:rp[u + 1]].astype(np.int64)
        ref.train_one_user(u, row[keep], negs)
        o.step_frozen([u], [row[keep]], [negs])
    assert_params_close(ref, o, 1e-10)


@pytest.mark.parametrize("kw", [dict(), dict(asymmetric=True), dict(corruption_ratio=1.0),
                                dict(tanh=True, user_factor=False)])
def test_recommend_matches_reference(orc, data, kw):
    cfg, ref, o = make_pair(orc, data, **kw)
    for u in range(data["U"]):
        ids, scores = o.recommend(u, 10)
        assert ids.tolist() == ref.recommend(u, 10).tolist()
        assert np.all(np.diff(scores) <= 0)


def test_recommend_tie_rule_keeps_lower_id(orc, data):
    """heap.hpp:44-52: a candidate replaces the minimum only if STRICTLY better; with all
    scores equal the first k unrated ids (ascending scan) survive."""
    cfg, ref, o = make_pair(orc, data)
    z = {k: np.zeros_like(v) for k, v in o.get_params().items() if not k.endswith("_ag")}
    ref.set_params(z)
    o.set_params(z)
    rp, col = data["train_row_ptr"], data["train_col"]
    for u in (0, 3):
        rated = set(col[rp[u]:rp[u + 1]].tolist())
        want = [i for i in range(data["I"]) if i not in rated][:10]
        assert sorted(ref.recommend(u, 10).tolist()) == want
        assert o.recommend(u, 10)[0].tolist() == want


def test_losses_and_representations(orc, data):
    for kw in (dict(corruption_ratio=0.0), dict(corruption_ratio=1.0, scaled=False),
               dict(corruption_ratio=0.0, asymmetric=True, loss="SQUARE")):
        cfg, ref, o = make_pair(orc, data, **kw)
        keep = None if cfg["corruption_ratio"] == 0.0 else np.zeros(len(data["train_col"]), np.uint8)
        assert abs(o.data_loss(keep) - ref.data_loss()) <= 1e-10 * abs(ref.data_loss())
        assert abs(o.penalty_loss() - ref.penalty_loss()) <= 1e-12 * abs(ref.penalty_loss())
        np.testing.assert_allclose(o.user_representations(), ref.user_representations(),
                                   rtol=1e-12, atol=1e-300)


def test_corruption_rule(orc, data):
    """cdae.hpp:366 keeps an item iff uniform() > ratio: ratio 1 keeps nothing, ratio 0
    keeps everything (up to a 2^-53 event), ratio .5 keeps about half, subset of the row."""
    cfg, ref, o = make_pair(orc, data)
    orc.Reference.seed(123, 123)
    rp, col = data["train_row_ptr"], data["train_col"]
    tot = kept = 0
    for u in range(data["U"]):
        row = set(col[rp[u]:rp[u + 1]].tolist())
        assert len(ref.corrupt(u, 1.0)) == 0
        assert set(ref.corrupt(u, 0.0).tolist()) == row
        k = ref.corrupt(u, 0.5)
        assert set(k.tolist()) <= row
        tot += len(row)
        kept += len(k)
    assert 0.4 < kept / tot < 0.6
    # the oracle's Philox mask obeys the same rule and rate
    kept = sum(int(o.sample_keep(1, 0, u).sum()) for u in range(data["U"]))
    assert 0.4 < kept / tot < 0.6


def test_topn_metrics(orc, data):
    rng = np.random.default_rng(2)
    for _ in range(50):
        lst = rng.permutation(40)[:10]
        test = rng.permutation(40)[:rng.integers(1, 15)]
        np.testing.assert_array_equal(orc.evaluate_rec_list(lst, test),
                                      orc.ref_evaluate_rec_list(lst, test))
    cfg, ref, o = make_pair(orc, data)
    got, n_users = o.topn_evaluate(data["test_row_ptr"], data["test_col"])
    want = ref.topn_evaluate(data["test_row_ptr"], data["test_col"])
    assert n_users == int(np.sum(np.diff(data["test_row_ptr"]) > 0))
    np.testing.assert_allclose(got, want, rtol=0, atol=6e-5)   # reference prints 5 significant digits


def test_negative_sampler_spec(orc, data):
    """recsys_model_base.hpp:46-57: never a positive of the user, always < I, n_u*num_neg draws."""
    cfg, ref, o = make_pair(orc, data)
    rp, col = data["train_row_ptr"], data["train_col"]
    seen = np.zeros(data["I"], np.int64)
    for u in range(data["U"]):
        negs = o.sample_negatives(99, 0, u)
        row = col[rp[u]:rp[u + 1]]
        assert len(negs) == len(row) * cfg["num_neg"]
        assert negs.min() >= 0 and negs.max() < data["I"]
        assert not np.isin(negs, row).any()
        seen += np.bincount(negs, minlength=data["I"])
    assert (seen > 0).mean() > 0.95       # roughly uniform coverage
