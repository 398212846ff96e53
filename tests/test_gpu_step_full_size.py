"""The training STEP (not just the forward pass) against the oracle at BASELINE.json's real sizes.

One frozen minibatch with explicit masks / negatives (cdae_train_users) is compared, row by row and
accumulator by accumulator, with Oracle.step_frozen — the frozen-batch restatement of
train_one_user_corruption (cdae.hpp:198-358) — at

  B        100,000 x 50,000, K=50, tied, AdaGrad: minibatch = users [24576, 32768) (8,192 users,
           the benchmark's batch size; ~1.5 M output rows, Zipf-head items receive ~10^3 fp32
           reductions each)
  D shard  125,000 x 200,000, K=100 (one of 8 ranks' users against the full item side)
  C        27,000 items, K=200, ~145 items per user, asymmetric, FULL-item decode (tcgen05 path):
           256 users = two 128-user tiles against every item
  E        100,000 items, K=256, full-item decode: 64 users

Sampled configurations: |a-b| <= 2e-5 + 2e-4*|b| (accumulators 1e-3 relative), the tolerance of
tests/test_gpu_parity.py.  Full decode: against the oracle's rounding=1 mode (bf16 operands, exact
products, fp64 sums), |a-b| <= 5e-4 + 1e-3*|b|.  The Zipf-head rows (most frequent items of the
minibatch) are additionally checked on their own so a failure names them.
"""
import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu

P_RTOL, P_ATOL, AG_RTOL = 2e-4, 2e-5, 1e-3


def _compare(m, o, names, rtol, atol, ag_rtol, head_items=None, ag_atol=None):
    worst = {}
    for k in names:
        a = m.get_param(k)
        if a.size == 0:
            continue
        b = np.asarray(o.param(k)).reshape(a.shape)
        rt = ag_rtol if k.endswith("_ag") else rtol
        at = ag_atol if (ag_atol is not None and k.endswith("_ag")) else atol
        err = np.abs(a - b) - (at + rt * np.abs(b))
        worst[k] = float(err.max())
        if head_items is not None and a.ndim == 2 and a.shape[0] == o.I:
            np.testing.assert_allclose(a[head_items], b[head_items], rtol=rt, atol=at,
                                       err_msg="Zipf-head rows of " + k)
        np.testing.assert_allclose(a, b, rtol=rt, atol=at, err_msg=k)
    return worst


@pytest.mark.parametrize("name,U,I,K,mean,u0,n", [
    ("B", 100_000, 50_000, 50, 30.0, 24_576, 8192),
    ("D-shard", 125_000, 200_000, 100, 30.0, 65_536, 8192),
])
def test_sampled_step_at_full_size(oracle_built, name, U, I, K, mean, u0, n):
    from cdae_b200 import CDAE, CDAEConfig, synth
    orc = oracle_built
    d = synth.make_dataset(U, I, mean, seed=20141119)
    rp, col = d["train_row_ptr"], d["train_col"]
    cfg = orc.default_config(loss="CE", num_dim=K, beta=1.0)              # tied, AdaGrad, q=.5 scaled, 5 negatives
    p = cases.random_params(U, I, K, 11, False, True)
    o = orc.Oracle(cfg, U, I, rp, col)
    o.set_params(p)
    m = CDAE(CDAEConfig(**cfg)).reset(U, I, rp, col)
    m.set_params(p)
    users = np.arange(u0, u0 + n)
    keeps = [o.sample_keep(5, 0, int(u)).astype(bool) for u in users]
    negs = [o.sample_negatives(5, 0, int(u)) for u in users]
    st = m.train_users(users, np.concatenate(keeps).astype(np.uint8), np.concatenate(negs).astype(np.int32))
    ls = o.step_frozen(users, [col[rp[u]:rp[u + 1]][k] for u, k in zip(users, keeps)], negs)
    n_out = int((rp[u0 + n] - rp[u0]) * (1 + cfg["num_neg"]))
    assert st.user_steps == n and st.outputs == n_out
    assert abs(st.loss_sum - ls) <= 1e-4 * abs(ls)
    # the most contended rows: items that are an output of the most users of this minibatch
    cnt = np.bincount(col[rp[u0]:rp[u0 + n]], minlength=I) + np.bincount(np.concatenate(negs), minlength=I)
    head = np.argsort(-cnt)[:32]
    assert cnt[head[0]] >= 500, cnt[head[0]]                              # really a Zipf head
    _compare(m, o, ["W", "W_ag", "b", "b_ag", "b_prime", "b_prime_ag", "Wu", "Wu_ag"], P_RTOL, P_ATOL, AG_RTOL, head)
    m.close()


@pytest.mark.parametrize("name,U,I,K,mean,n", [
    ("C", 4096, 27_000, 200, 145.0, 256),
    ("E", 2048, 100_000, 256, 50.0, 64),
])
def test_full_decode_step_at_full_item_count(oracle_built, name, U, I, K, mean, n):
    from cdae_b200 import CDAE, CDAEConfig, synth
    orc = oracle_built
    d = synth.make_dataset(U, I, mean, seed=20141119)
    rp, col = d["train_row_ptr"], d["train_col"]
    cfg = orc.default_config(loss="CE", num_dim=K, beta=1.0, asymmetric=True)
    p = cases.random_params(U, I, K, 12, True, True)
    o = orc.Oracle(cfg, U, I, rp, col)
    o.set_params(p)
    m = CDAE(CDAEConfig(full_decode=True, **cfg)).reset(U, I, rp, col)
    m.set_params(p)
    users = np.arange(100, 100 + n)
    keeps = [o.sample_keep(5, 0, int(u)).astype(bool) for u in users]
    st = m.train_users(users, np.concatenate(keeps).astype(np.uint8), None)
    ls = o.step_frozen_full(users, [col[rp[u]:rp[u + 1]][k] for u, k in zip(users, keeps)], rounding=1)
    assert st.user_steps == n and st.outputs == n * I
    assert np.isfinite(ls)          # (the tcgen05 path does not evaluate the loss: stats.loss_sum stays 0, include/cdae_b200.h)
    cnt = np.bincount(col[rp[100]:rp[100 + n]], minlength=I)
    head = np.argsort(-cnt)[:32]
    _compare(m, o, ["W", "V", "V_ag", "b", "b_prime", "b_prime_ag", "Wu"], 1e-3, 5e-4, 4e-3, head, ag_atol=1e-2)   # accumulators: sums of squared sums (test_gpu_fulldec.py)
    m.close()
