"""Full-item-decode training (output set = all items, three tcgen05 contractions)
against the oracle, through the C ABI.

The reference has no such function; the oracle's full mode is pinned to its frozen-batch mode
("negatives = every non-positive once", tests/test_oracle_golden.py), which is pinned to the
verbatim reference.  Two comparisons:
  * rounding=1 oracle (restates the bf16 rounding of z, W', g where they enter the contractions):
    what is left is fp32 accumulation order, the sigmoid polynomial (<2e-6) and rare one-ulp bf16
    flips of g -> |a-b| <= 5e-4 + 1e-3*|b| (measured: 2e-5 .. 2.4e-4; the sums run in atomics order), accumulators 1e-2 + 4e-3*|b|.
  * rounding=0 oracle (plain fp64): bounds the effect of bf16 operands themselves on one step ->
    |a-b| <= 3e-2 + 3e-2*|b| on parameters (accumulators, sums of squared sums, are only compared
    with the rounding=1 oracle).
"""
import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu

CASES = [
    # U, I, K, cfg overrides
    (300, 1000, 50, dict(loss="CE", beta=1.0)),                                   # tied, 1 k-block
    (200, 700, 100, dict(loss="SQUARE", asymmetric=True, learn_rate=0.01)),        # 2 k-blocks
    (260, 900, 200, dict(loss="CE", asymmetric=True, beta=1.0)),                   # config C's K, 4 k-blocks
    (150, 520, 130, dict(loss="CE", using_adagrad=False, learn_rate=0.01)),        # 3 k-blocks, plain SGD
    # K+2 = 64 exactly.  beta = 1 (the app's default, yelp.cpp:47): with beta = 0 and accumulators
    # near 1e-4 a step is lr*g/sqrt(acc), which amplifies a single bf16 rounding flip of one z
    # element (fp32 vs fp64 hidden value on either side of a bf16 boundary) by up to 10x
    (140, 600, 62, dict(loss="CE", asymmetric=True, linear_function=True, tanh=True, beta=1.0)),
    # no room for the two bias columns (K + 2 > round_up(K, 64)): b' is added in the score epilogue and
    # its gradient comes from the ones-operand MMA of the item-gradient kernel
    (200, 800, 256, dict(loss="CE", asymmetric=True, beta=1.0)),                    # config E's K
    (130, 500, 64, dict(loss="SQUARE", beta=1.0, learn_rate=0.01)),                 # tied
    (150, 640, 63, dict(loss="CE", asymmetric=True, beta=1.0)),
]


def build(orc, U, I, K, kw, seed=5, batch_users=0):
    from cdae_b200 import CDAE, CDAEConfig
    cfg = orc.default_config(num_dim=K, **kw)
    data = cases.small_dataset(U=U, I=I, mean=14.0, seed=seed)
    U, I = data["U"], data["I"]
    rp, col = data["train_row_ptr"], data["train_col"]
    p = cases.random_params(U, I, K, seed, cfg["asymmetric"], cfg["user_factor"], cfg["linear_function"])
    mk = dict(cfg)
    mk.update(full_decode=True, batch_users=batch_users)
    m = CDAE(CDAEConfig(**mk)).reset(U, I, rp, col)
    m.set_params(p)
    oracles = []
    for _ in range(2):
        o = orc.Oracle(cfg, U, I, rp, col)
        o.set_params(p)
        oracles.append(o)
    return cfg, data, m, oracles


AG_ATOL = 1e-2   # accumulators are sums of SQUARED gradient sums: |d acc| ~ 2|g||dg|, absolute in g's scale


def compare(m, o, rtol, atol, ag_rtol, what):
    from cdae_b200._lib import PARAMS
    bad = []
    for k in PARAMS:
        if ag_rtol is None and k.endswith("_ag"):
            continue
        a = m.get_param(k)
        if a.size == 0:
            continue
        b = np.asarray(o.param(k)).reshape(a.shape)
        r = ag_rtol if k.endswith("_ag") else rtol
        at = AG_ATOL if k.endswith("_ag") else atol
        err = np.abs(a - b) - (at + r * np.abs(b))
        print("%s %-10s max|a-b| %.3e  max|b| %.3e  worst excess %.3e" % (what, k, np.abs(a - b).max(), np.abs(b).max(), err.max()))
        if err.max() > 0:
            bad.append((k, float(np.abs(a - b).max())))
    return bad


@pytest.mark.parametrize("U,I,K,kw", CASES)
def test_fulldec_minibatch(oracle_built, U, I, K, kw):
    orc = oracle_built
    cfg, data, m, (o_bf, o_64) = build(orc, U, I, K, kw)
    rp, col = data["train_row_ptr"], data["train_col"]
    rng = np.random.default_rng(11)
    users = np.arange(data["U"])
    keep = (rng.random(len(col)) > cfg["corruption_ratio"]).astype(np.uint8)
    ins = [col[rp[u]:rp[u + 1]][keep[rp[u]:rp[u + 1]].astype(bool)] for u in users]
    st = m.train_users(users, keep, None)
    assert st.user_steps == data["U"] and st.outputs == data["U"] * data["I"]
    o_bf.step_frozen_full(users, ins, rounding=1)
    o_64.step_frozen_full(users, ins, rounding=0)
    bad = compare(m, o_bf, 1e-3, 5e-4, 4e-3, "bf16-oracle")
    bad64 = compare(m, o_64, 3e-2, 3e-2, None, "fp64-oracle")
    assert not bad, bad
    assert not bad64, bad64


def test_fulldec_epoch_minibatches(oracle_built):
    """train_one_iteration: Philox masks, three frozen minibatches (the last one partial)."""
    orc = oracle_built
    cfg, data, m, (o_bf, _) = build(orc, 300, 800, 50, dict(loss="CE", beta=1.0, asymmetric=True), batch_users=128)
    for ep in range(2):
        st = m.train_one_iteration(seed=17, epoch=ep)
        o_bf.train_epoch_full(17, ep, 128, rounding=1)
        assert st.user_steps == data["U"]
    bad = compare(m, o_bf, 2e-3, 2e-4, 8e-3, "epoch")
    assert not bad, bad
    ids, _ = m.recommend_all(10)
    same = sum(ids[u].tolist() == o_bf.recommend(u, 10)[0].tolist() for u in range(data["U"]))
    assert same >= 0.9 * data["U"], same      # lists come from slightly different parameters


def test_fulldec_rejects_unsupported(oracle_built):
    from cdae_b200 import CDAE, CDAEConfig, CdaeError
    data = cases.small_dataset(U=40, I=100, mean=8.0, seed=2)
    for kw in (dict(loss="HINGE", num_dim=10), dict(loss="CE", num_dim=300)):
        with pytest.raises(CdaeError):
            CDAE(CDAEConfig(full_decode=True, **kw)).reset(data["U"], data["I"], data["train_row_ptr"], data["train_col"])
