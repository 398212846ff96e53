"""Parity of the CUDA path (through the C ABI) with the oracle and the golden vectors.

Tolerances: the device computes in fp32, the reference in fp64.
  hidden activations   : max relative error <= 1e-4 (north-star bound; measured ~1e-6)
  parameters after step: |a-b| <= 2e-5 + 2e-4*|b|  (fp32 rounding + atomics order);
                         AdaGrad accumulators 2e-5 + 1e-3*|b|
  top-N lists          : identical ids
"""
import numpy as np
import pytest

from tests import cases, golden

pytestmark = pytest.mark.gpu

Z_RTOL = 1e-4
P_RTOL, P_ATOL = 2e-4, 2e-5
AG_RTOL = 1e-3


@pytest.fixture(scope="module")
def orc(oracle_built):
    return oracle_built


def gpu_model(cfg, U, I, rp, col, params, batch_users=0):
    from cdae_b200 import CDAE, CDAEConfig
    kw = dict(cfg)
    kw["batch_users"] = batch_users
    m = CDAE(CDAEConfig(**kw)).reset(U, I, rp, col)
    m.set_params(params)
    return m


def assert_params(m, o, names=None):
    from cdae_b200._lib import PARAMS
    for k in names or PARAMS:
        a = m.get_param(k)
        if a.size == 0:
            continue
        b = np.asarray(o[k] if isinstance(o, dict) else o.param(k)).reshape(a.shape)
        # AdaGrad accumulators hold sums of SQUARED summed gradients: twice the relative error
        # of an fp32 sum taken in atomics order, so they get a wider relative band
        rtol = AG_RTOL if k.endswith("_ag") else P_RTOL
        np.testing.assert_allclose(a, b, rtol=rtol, atol=P_ATOL, err_msg=k)


@pytest.mark.parametrize("name", golden.NAMES)
def test_golden_forward(name):
    g = golden.load(name)
    cfg = g["cfg"]
    m = gpu_model(cfg, g["U"], g["I"], g["train_row_ptr"], g["train_col"], g["p0"])
    scale = 1.0 / (1.0 - cfg["corruption_ratio"]) if cfg["scaled"] else 1.0
    users = np.arange(g["U"])
    z = m.encode(users)
    zc = m.encode(users, g["keep"], scale)
    for got, want in ((z, g["z_clean"]), (zc, g["z_corrupt"])):
        err = np.abs(got - want) / np.maximum(np.abs(want), 1e-3)
        assert err.max() <= Z_RTOL, err.max()
    ids, _ = m.recommend_all(10)
    assert ids.tolist() == g["rec_before"].tolist()
    assert abs(m.penalty_loss() - g["penalty_before"]) <= 1e-5 * abs(g["penalty_before"])
    if not np.isnan(g["data_loss_q0"]):
        assert abs(m.data_loss() - g["data_loss_q0"]) <= 1e-4 * abs(g["data_loss_q0"])
    met, n = m.topn_evaluate(g["test_row_ptr"], g["test_col"])
    np.testing.assert_allclose(met, g["metrics_before"], atol=6e-5)


@pytest.mark.parametrize("name", golden.NAMES)
def test_golden_online_pass(name):
    """batch of one user at a time == the reference's train_one_user_corruption sequence."""
    g = golden.load(name)
    cfg, rp = g["cfg"], g["train_row_ptr"]
    nu = cfg["num_neg"]
    m = gpu_model(cfg, g["U"], g["I"], rp, g["train_col"], g["p0"], batch_users=1)
    for u in range(g["U"]):
        m.train_users([u], g["keep"][rp[u]:rp[u + 1]], g["negs"][rp[u] * nu:rp[u + 1] * nu])
    assert_params(m, g["p1"], list(g["p1"].keys()))
    ids, _ = m.recommend_all(10)
    same = sum(a == b for a, b in zip(ids.tolist(), g["rec_after"].tolist()))
    assert same == g["U"], "%d of %d top-10 lists identical" % (same, g["U"])


@pytest.mark.parametrize("kw", cases.CONFIG_GRID)
def test_frozen_minibatch_matches_oracle(orc, kw):
    kw = dict(kw)
    kw.setdefault("loss", "CE")
    cfg = orc.default_config(**kw)
    data = cases.small_dataset(U=96, I=300, mean=14.0, seed=21)
    U, I, K = data["U"], data["I"], cfg["num_dim"]
    rp, col = data["train_row_ptr"], data["train_col"]
    p = cases.random_params(U, I, K, 5, cfg["asymmetric"], cfg["user_factor"], cfg["linear_function"])
    o = orc.Oracle(cfg, U, I, rp, col)
    o.set_params(p)
    m = gpu_model(cfg, U, I, rp, col, p)
    rng = np.random.default_rng(17)
    steps = cases.draw_step_inputs(data, cfg["num_neg"], cfg["corruption_ratio"], rng)
    for users in (np.arange(0, 64), np.arange(64, 96), np.array([3, 90, 41])):
        keep = np.concatenate([steps[u][0] for u in users]).astype(np.uint8)
        negs = np.concatenate([steps[u][1] for u in users] + [np.zeros(0, np.int64)]).astype(np.int32)
        st = m.train_users(users, keep, negs)
        ls = o.step_frozen(users, [col[rp[u]:rp[u + 1]][steps[u][0]] for u in users],
                           [steps[u][1] for u in users])
        assert st.user_steps == len(users)
        assert st.outputs == sum((rp[u + 1] - rp[u]) * (1 + cfg["num_neg"]) for u in users)
        assert abs(st.loss_sum - ls) <= 1e-4 * max(1.0, abs(ls))
    assert_params(m, o)


@pytest.mark.parametrize("kw", [dict(loss="CE", beta=1.0), dict(loss="SQUARE", asymmetric=True),
                                dict(loss="CE", num_corruptions=2, corruption_ratio=0.3),
                                dict(loss="CE", corruption_ratio=0.0, scaled=False),
                                dict(loss="CE", corruption_ratio=1.0, scaled=False),
                                dict(loss="CE", corruption_ratio=1.0, beta=1.0)])   # q = 1 AND scaled: no inf * 0
def test_epoch_with_device_sampling_matches_oracle(orc, kw):
    """cdae_train_epoch: Philox masks/negatives on the device == the same streams on the CPU."""
    cfg = orc.default_config(**kw)
    data = cases.small_dataset(U=200, I=400, mean=12.0, seed=31)
    U, I, K = data["U"], data["I"], cfg["num_dim"]
    rp, col = data["train_row_ptr"], data["train_col"]
    p = cases.random_params(U, I, K, 6, cfg["asymmetric"], cfg["user_factor"], warm=False)
    o = orc.Oracle(cfg, U, I, rp, col)
    o.set_params(p)
    m = gpu_model(cfg, U, I, rp, col, p, batch_users=48)
    kept = sum(int(o.sample_keep(77, 0, u).sum()) for u in range(U)) if cfg["num_corruptions"] == 1 else None
    for epoch in range(2):
        st = m.train_one_iteration(seed=77, epoch=epoch)
        ls = o.train_epoch(77, epoch, batch_users=48)
        assert st.user_steps == U * cfg["num_corruptions"]
        assert st.outputs == len(col) * (1 + cfg["num_neg"]) * cfg["num_corruptions"]
        if epoch == 0 and kept is not None:
            assert st.inputs_kept == kept
        assert abs(st.loss_sum - ls) <= 2e-4 * abs(ls)
    assert_params(m, o)
    assert abs(m.penalty_loss() - o.penalty_loss()) <= 1e-4 * o.penalty_loss()
    # host-CSR form of the same call
    m2 = gpu_model(cfg, U, I, rp, col, p, batch_users=48)
    o2 = orc.Oracle(cfg, U, I, rp, col)
    o2.set_params(p)
    m2.train_one_iteration(seed=77, epoch=0, csr=(rp, col))
    o2.train_epoch(77, 0, batch_users=48)
    assert_params(m2, o2)


def test_device_init_matches_oracle_stream(orc):
    cfg = orc.default_config(loss="CE", asymmetric=True, num_dim=50)
    data = cases.small_dataset(U=64, I=200, mean=10.0, seed=41)
    o = orc.Oracle(cfg, data["U"], data["I"], data["train_row_ptr"], data["train_col"])
    o.init_params(123)
    m = gpu_model(cfg, data["U"], data["I"], data["train_row_ptr"], data["train_col"], {})
    m.init_params(123)
    for k in ("W", "V", "Wu"):
        np.testing.assert_array_equal(m.get_param(k), o.param(k))
        s = 4 * np.sqrt(6.0 / (data["I"] + 50))
        assert np.abs(m.get_param(k)).max() <= s and np.abs(m.get_param(k)).max() > 0.9 * s
    for k in ("W_ag", "b_prime_ag"):
        np.testing.assert_allclose(m.get_param(k), 1e-4, rtol=1e-6)
    assert np.all(m.get_param("b") == 0) and np.all(m.get_param("b_prime") == 0)


@pytest.mark.parametrize("K", [3, 10, 50, 100, 200, 256, 300])
def test_every_row_geometry(orc, K):
    """One frozen step + top-N for each <lanes-per-row, vectors-per-lane> instantiation."""
    cfg = orc.default_config(loss="CE", num_dim=K, num_neg=2)
    data = cases.small_dataset(U=40, I=150, mean=9.0, seed=51)
    U, I = data["U"], data["I"]
    rp, col = data["train_row_ptr"], data["train_col"]
    p = cases.random_params(U, I, K, 8, False, True)
    o = orc.Oracle(cfg, U, I, rp, col)
    o.set_params(p)
    m = gpu_model(cfg, U, I, rp, col, p)
    rng = np.random.default_rng(3)
    steps = cases.draw_step_inputs(data, 2, 0.5, rng)
    users = np.arange(U)
    keep = np.concatenate([steps[u][0] for u in users]).astype(np.uint8)
    negs = np.concatenate([steps[u][1] for u in users]).astype(np.int32)
    z = m.encode(users, keep, 2.0)
    zo = np.stack([o.hidden(u, col[rp[u]:rp[u + 1]][steps[u][0]], 2.0) for u in users])
    assert (np.abs(z - zo) / np.maximum(np.abs(zo), 1e-3)).max() <= Z_RTOL
    m.train_users(users, keep, negs)
    o.step_frozen(users, [col[rp[u]:rp[u + 1]][steps[u][0]] for u in users], [steps[u][1] for u in users])
    assert_params(m, o)
    ids, _ = m.recommend_all(10)
    for u in users:
        assert ids[u].tolist() == o.recommend(u, 10)[0].tolist()


def test_ragged_and_edge_rows(orc):
    """Users with 1 item, a user with hundreds of items (many chunks), num_neg = 0."""
    rng = np.random.default_rng(9)
    I = 700
    lens = [1, 1, 2, 65, 129, 400, 3, 17]
    rows = [np.sort(rng.choice(I, n, replace=False)) for n in lens]
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    col = np.concatenate(rows).astype(np.int32)
    U = len(lens)
    for nu in (0, 1, 5):
        cfg = orc.default_config(loss="CE", num_neg=nu, num_dim=20)
        p = cases.random_params(U, I, 20, 4, False, True)
        o = orc.Oracle(cfg, U, I, rp, col)
        o.set_params(p)
        m = gpu_model(cfg, U, I, rp, col, p)
        keep = (rng.random(len(col)) > 0.5).astype(np.uint8)
        negs = [rng.choice(np.setdiff1d(np.arange(I), rows[u]), lens[u] * nu) for u in range(U)]
        m.train_users(np.arange(U), keep, np.concatenate(negs + [np.zeros(0, np.int64)]).astype(np.int32))
        o.step_frozen(np.arange(U), [rows[u][keep[rp[u]:rp[u + 1]].astype(bool)] for u in range(U)], negs)
        assert_params(m, o)
        st = m.train_one_iteration(seed=5, epoch=0)
        o.train_epoch(5, 0, batch_users=8192)
        assert st.user_steps == U
        assert_params(m, o)


def test_recommend_and_metrics_match_oracle(orc):
    for kw in (dict(loss="CE"), dict(loss="SQUARE", asymmetric=True, num_dim=50),
               dict(loss="CE", corruption_ratio=1.0, scaled=False)):
        cfg = orc.default_config(**kw)
        data = cases.small_dataset(U=150, I=500, mean=12.0, seed=61)
        U, I, K = data["U"], data["I"], cfg["num_dim"]
        p = cases.random_params(U, I, K, 2, cfg["asymmetric"], cfg["user_factor"])
        o = orc.Oracle(cfg, U, I, data["train_row_ptr"], data["train_col"])
        o.set_params(p)
        m = gpu_model(cfg, U, I, data["train_row_ptr"], data["train_col"], p)
        ids, sc = m.recommend_all(10)
        for u in range(U):
            oi, os_ = o.recommend(u, 10)
            assert ids[u].tolist() == oi.tolist()
            np.testing.assert_allclose(sc[u], os_, rtol=1e-4, atol=1e-5)
            li, _ = m.recommend(u, 10)
            assert li.tolist() == oi.tolist()
        met, n = m.topn_evaluate(data["test_row_ptr"], data["test_col"])
        want, n2 = o.topn_evaluate(data["test_row_ptr"], data["test_col"])
        assert n == n2
        np.testing.assert_allclose(met, want, atol=1e-12)


def test_tie_rule_and_short_candidate_error(orc):
    """All-zero weights: every score ties, the first k unrated ids must come back
    (heap.hpp:44-52).  Fewer than k unrated items is an error (CHECK_EQ cdae.hpp:187)."""
    from cdae_b200 import CdaeError
    cfg = orc.default_config(loss="CE")
    data = cases.small_dataset(U=30, I=120, mean=9.0, seed=71)
    m = gpu_model(cfg, data["U"], data["I"], data["train_row_ptr"], data["train_col"], {})
    ids, _ = m.recommend_all(10)
    rp, col = data["train_row_ptr"], data["train_col"]
    for u in range(data["U"]):
        rated = set(col[rp[u]:rp[u + 1]].tolist())
        assert ids[u].tolist() == [i for i in range(data["I"]) if i not in rated][:10]
    rp2 = np.array([0, 8], np.int64)
    col2 = np.arange(8, dtype=np.int32)
    m2 = gpu_model(cfg, 1, 12, rp2, col2, {})
    with pytest.raises(CdaeError):
        m2.pre_recommend(10)


def test_data_loss_matches_oracle(orc):
    for kw in (dict(loss="CE"), dict(loss="SQUARE", num_corruptions=3, corruption_ratio=0.4)):
        cfg = orc.default_config(**kw)
        data = cases.small_dataset(U=120, I=300, mean=12.0, seed=81)
        U, I, K = data["U"], data["I"], cfg["num_dim"]
        rp, col = data["train_row_ptr"], data["train_col"]
        p = cases.random_params(U, I, K, 2, False, True)
        o = orc.Oracle(cfg, U, I, rp, col)
        o.set_params(p)
        m = gpu_model(cfg, U, I, rp, col, p)
        want = 0.0
        for c in range(cfg["num_corruptions"]):
            keep = np.concatenate([o.sample_keep(9, 0x80000000 + c, u) for u in range(U)])
            want += o.data_loss(keep)
        want /= cfg["num_corruptions"]
        assert abs(m.data_loss(seed=9) - want) <= 1e-4 * abs(want)
        assert abs(m.current_loss(seed=9) - (want + o.penalty_loss())) <= 1e-4 * abs(want)


def test_errors_are_reported_not_swallowed(orc):
    from cdae_b200 import CDAE, CDAEConfig, CdaeError
    data = cases.small_dataset(U=20, I=100, mean=9.0, seed=91)
    rp, col = data["train_row_ptr"], data["train_col"]
    with pytest.raises(CdaeError):                       # unsorted row
        bad = col.copy()
        bad[0], bad[1] = bad[1], bad[0]
        CDAE(CDAEConfig(loss="CE")).reset(data["U"], data["I"], rp, bad)
    with pytest.raises(CdaeError):                       # item id out of range
        CDAE(CDAEConfig(loss="CE")).reset(data["U"], int(col.max()), rp, col)
    m = CDAE(CDAEConfig(loss="CE")).reset(data["U"], data["I"], rp, col)
    with pytest.raises(CdaeError):                       # duplicate user in a frozen batch
        n = rp[1] - rp[0]
        m.train_users([0, 0], np.ones(2 * n, np.uint8), np.zeros(2 * n * 5, np.int32) + int(col.max()))
    # the struct-default LOGISTIC loss aborts in the reference: here it is an error
    m = CDAE(CDAEConfig()).reset(data["U"], data["I"], rp, col)
    m.init_params(1)
    with pytest.raises(CdaeError):
        m.train_one_iteration(seed=1, epoch=0)


def test_checkpoint_round_trip(orc, tmp_path):
    """cdae_save / cdae_load: every block incl. AdaGrad state survives, a model restored
    into a fresh handle scores and recommends identically, and mismatching shapes are refused."""
    from cdae_b200 import CdaeError
    from cdae_b200._lib import PARAMS
    cfg = orc.default_config(loss="CE", num_dim=20, beta=1.0, asymmetric=True, linear_function=True)
    data = cases.small_dataset(U=90, I=300, mean=10.0, seed=4)
    U, I, rp, col = data["U"], data["I"], data["train_row_ptr"], data["train_col"]
    m = gpu_model(cfg, U, I, rp, col, cases.random_params(U, I, 20, 2, True, True, True))
    m.train_one_iteration(seed=3, epoch=0)
    path = tmp_path / "model.ckpt"
    m.save(path)
    m2 = gpu_model(cfg, U, I, rp, col, {})
    m2.load(path)
    for k in PARAMS:
        np.testing.assert_array_equal(m.get_param(k), m2.get_param(k), err_msg=k)
    assert m.recommend_all(10)[0].tolist() == m2.recommend_all(10)[0].tolist()
    m3 = gpu_model(orc.default_config(loss="CE", num_dim=21, beta=1.0, asymmetric=True, linear_function=True),
                   U, I, rp, col, {})
    with pytest.raises(CdaeError):
        m3.load(path)


def test_q1_scaled_losses_are_finite(orc):
    """q == 1 with scaled made scale = inf and every hidden value NaN."""
    cfg = orc.default_config(loss="CE", corruption_ratio=1.0, beta=1.0)
    data = cases.small_dataset(U=64, I=200, mean=10.0, seed=4)
    U, I = data["U"], data["I"]
    rp, col = data["train_row_ptr"], data["train_col"]
    p = cases.random_params(U, I, cfg["num_dim"], 3, False, True)
    o = orc.Oracle(cfg, U, I, rp, col)
    o.set_params(p)
    m = gpu_model(cfg, U, I, rp, col, p)
    dl = m.data_loss(seed=1)
    assert np.isfinite(dl)
    assert abs(dl - o.data_loss(np.zeros(len(col), np.uint8))) <= 1e-4 * abs(dl)
    ids, _ = m.recommend_all(10)           # cdae.hpp:168: the input set is empty when q == 1
    for u in range(U):
        assert ids[u].tolist() == o.recommend(u, 10)[0].tolist()


def test_train_epoch_csr_rejects_a_bad_csr(orc):
    """the host-CSR call must not index tables with unchecked ids.  The column array is checked on
    the device, minibatch by minibatch as its rows arrive; a violation is an error, the offending minibatch and
    all later ones leave the parameters as they were, and training is refused until a valid CSR arrives."""
    from cdae_b200 import CdaeError
    cfg = orc.default_config(loss="CE", beta=1.0)
    data = cases.small_dataset(U=150, I=300, mean=10.0, seed=6)
    U, I = data["U"], data["I"]
    rp, col = data["train_row_ptr"], data["train_col"]
    p = cases.random_params(U, I, cfg["num_dim"], 3, False, True)
    for kind in ("range", "negative", "unsorted"):
        m = gpu_model(cfg, U, I, rp, col, p, batch_users=64)
        before = m.get_params()
        bad = col.copy()
        s = int(rp[70]) + 1
        if kind == "range":
            bad[s] = I + 12345
        elif kind == "negative":
            bad[s] = -7
        else:
            bad[s - 1], bad[s] = bad[s], bad[s - 1]
        with pytest.raises(CdaeError):
            m.train_one_iteration(seed=3, epoch=0, csr=(rp, bad))
        after = m.get_params()
        # the upload is pipelined per minibatch (64 users here): the minibatch before the offending one (users
        # 0..63) has been trained, the offending one (user 70) and every later one have not touched anything
        assert np.array_equal(before["Wu"][64:], after["Wu"][64:]), kind
        assert not np.array_equal(before["Wu"][:64], after["Wu"][:64]), kind
        o = orc.Oracle(cfg, U, I, rp, col)
        o.set_params(p)
        o.train_epoch(3, 0, batch_users=64, u0=0, u1=64)
        for k in ("W", "b", "b_prime", "Wu"):
            np.testing.assert_allclose(after[k], o.param(k).reshape(after[k].shape), rtol=P_RTOL, atol=P_ATOL, err_msg=kind + " " + k)
        with pytest.raises(CdaeError):
            m.train_one_iteration(seed=3, epoch=0)                       # resident CSR is known bad
        # a valid CSR heals the handle and trains exactly like a fresh one
        m.set_params(p)
        m.train_one_iteration(seed=3, epoch=0, csr=(rp, col))
        m2 = gpu_model(cfg, U, I, rp, col, p, batch_users=64)
        m2.train_one_iteration(seed=3, epoch=0)
        for k in ("W", "b", "b_prime", "Wu"):
            np.testing.assert_allclose(m.get_param(k), m2.get_param(k), rtol=1e-5, atol=1e-6, err_msg=kind + " " + k)
    # dtype coercion in the Python mirror: an int32 row_ptr is converted, not reinterpreted
    m = gpu_model(cfg, U, I, rp, col, p, batch_users=64)
    m.train_one_iteration(seed=3, epoch=0, csr=(rp.astype(np.int32), col.astype(np.int64)))


def test_c_abi_refuses_stale_recommendations(orc):
    """cdae_topn_lookup / fetch / evaluate must not serve lists built before a parameter
    update — checked at the C ABI, below the wrappers' own bookkeeping."""
    import ctypes as C
    from cdae_b200 import _lib
    cfg = orc.default_config(loss="CE", beta=1.0)
    data = cases.small_dataset(U=64, I=200, mean=10.0, seed=8)
    U, I = data["U"], data["I"]
    rp, col = data["train_row_ptr"], data["train_col"]
    p = cases.random_params(U, I, cfg["num_dim"], 3, False, True)
    m = gpu_model(cfg, U, I, rp, col, p)
    L = _lib.lib()
    ids = np.zeros(10, np.int64)

    def lookup():
        return L.cdae_topn_lookup(m._h, 3, ids.ctypes.data_as(_lib.i64p), None)

    assert L.cdae_topn_build(m._h, 10) == 0 and lookup() == 0
    m.train_one_iteration(seed=1, epoch=0)
    assert lookup() == -4                                                # CDAE_E_STATE
    assert L.cdae_topn_build(m._h, 10) == 0 and lookup() == 0
    m.set_params({"b": p["b"]})
    assert lookup() == -4
    assert L.cdae_topn_build(m._h, 10) == 0 and lookup() == 0
    m.init_params(5)
    assert lookup() == -4
    # unsorted test rows are refused by cdae_topn_evaluate (it binary-searches them)
    assert L.cdae_topn_build(m._h, 10) == 0
    trp, tcol = data["test_row_ptr"], data["test_col"].copy()
    u = int(np.argmax(np.diff(trp) >= 2))
    tcol[trp[u]], tcol[trp[u] + 1] = tcol[trp[u] + 1], tcol[trp[u]]
    with pytest.raises(_lib.CdaeError):
        m.topn_evaluate(trp, tcol)
