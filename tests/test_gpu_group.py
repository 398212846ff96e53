"""Single-process multi-GPU mode (cdae_group_*, csrc/group.inl): one process drives N GPUs — what lets the
reference's one-process app (Solver<CDAE>::train, solver-inl.hpp:19,53,55) use the box.  Same parity bar
as the process-group tests: the oracle's frozen-batch epochs on the same global minibatches.  Needs >= 2
GPUs (skipped on the driver's 1-GPU box; `gpurun --gpus 2` runs it)."""
import os

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


def _need_gpus(n):
    import torch
    if torch.cuda.device_count() < n:
        pytest.skip("needs >= %d GPUs" % n)


@pytest.mark.parametrize("kw", [dict(loss="CE", beta=1.0, num_dim=50), dict(loss="SQUARE", asymmetric=True, num_dim=20),
                                dict(loss="CE", beta=1.0, asymmetric=True, num_dim=100, full_decode=True)])
def test_group_training_matches_oracle(oracle_built, kw, tmp_path):
    _need_gpus(2)
    import torch
    from cdae_b200 import CDAEGroup, CDAEConfig
    orc = oracle_built
    kw = dict(kw)
    full = kw.pop("full_decode", False)
    n_gpu = 4 if torch.cuda.device_count() >= 4 else 2
    cfg = orc.default_config(**kw)
    data = cases.small_dataset(U=403, I=500, mean=12.0, seed=5)
    U, I, K = data["U"], data["I"], cfg["num_dim"]
    rp, col = data["train_row_ptr"], data["train_col"]
    p = cases.random_params(U, I, K, 9, cfg["asymmetric"], cfg["user_factor"])
    B = 96
    m = CDAEGroup(CDAEConfig(batch_users=B, full_decode=full, **cfg), devices=list(range(n_gpu))).reset(U, I, rp, col)
    m.set_params(p)
    o = orc.Oracle(cfg, U, I, rp, col)
    o.set_params(p)
    for epoch in range(3):
        st = m.train_one_iteration(seed=123, epoch=epoch, csr=(rp, col) if epoch == 2 else None)
        assert st.user_steps == U
        if full:
            o.train_epoch_full(123, epoch, B, rounding=1)
        else:
            o.train_epoch(123, epoch, batch_users=B)
    ex_users = np.arange(7, 7 + 81)
    ex = cases.draw_step_inputs(data, cfg["num_neg"], cfg["corruption_ratio"], np.random.default_rng(77), ex_users)
    m.train_users(ex_users, np.concatenate([ex[u][0] for u in ex_users]).astype(np.uint8),
                  None if full else np.concatenate([ex[u][1] for u in ex_users]).astype(np.int32))
    ins = [col[rp[u]:rp[u + 1]][ex[u][0]] for u in ex_users]
    if full:
        o.step_frozen_full(ex_users, ins, rounding=1)
    else:
        o.step_frozen(ex_users, ins, [ex[u][1] for u in ex_users])
    for k in ("W", "V", "Wu", "b", "b_prime", "W_ag", "V_ag", "Wu_ag", "b_ag"):
        a = m.get_param(k)
        if a.size == 0:
            continue
        ref = o.param(k)
        err = np.abs(a - ref).max() / max(1e-12, np.abs(ref).max())
        assert err <= ((8e-3 if k.endswith("_ag") else 3e-3) if full else 2e-4), (k, err)
    assert abs(m.penalty_loss() - o.penalty_loss()) <= 1e-4 * o.penalty_loss() * (30 if full else 1)
    keep = np.concatenate([o.sample_keep(7, 0x80000000, u) for u in range(U)])
    assert abs(m.data_loss(seed=7) - o.data_loss(keep)) <= (2e-3 if full else 2e-4) * abs(o.data_loss(keep))
    # lists and hidden vectors are served by the GPU that owns the user
    o2 = orc.Oracle(cfg, U, I, rp, col)
    o2.set_params({k: v for k, v in m.get_params().items() if v.size})
    users = np.arange(0, U, 9)
    z = m.encode(users)
    for j, u in enumerate(users):
        assert m.recommend(int(u), 10)[0].tolist() == o2.recommend(int(u), 10)[0].tolist(), u
        np.testing.assert_allclose(z[j], o2.hidden(int(u), col[rp[u]:rp[u + 1]].astype(np.int64), 1.0), rtol=1e-4, atol=1e-6)
    # checkpoint round trip through the group
    path = str(tmp_path / "ckpt.bin")
    before = {k: m.get_param(k) for k in ("W", "Wu", "b", "W_ag")}
    m.save(path)
    m.set_params({"b": np.zeros(K)})
    m.load(path)
    for k, v in before.items():
        assert np.array_equal(m.get_param(k), v), k
    m.close()


def test_yelp_app_on_two_gpus(tmp_path):
    """The unchanged reference app on libcf::CDAE with CDAE_B200_GPUS=2 trains like the 1-GPU run."""
    _need_gpus(2)
    from tests.test_host_app import B200, _write_pairs, _run_app
    if not os.path.exists(B200):
        pytest.skip("prebuilt app binary missing")
    runs = {}
    for name, env in (("one", {"CDAE_B200_BATCH_USERS": "64", "CDAE_B200_SEED": "7"}),
                      ("two", {"CDAE_B200_BATCH_USERS": "64", "CDAE_B200_SEED": "7", "CDAE_B200_GPUS": "2"})):
        d = tmp_path / name
        d.mkdir()
        _write_pairs(str(d / "yelp_10core.txt"))
        runs[name] = _run_app(B200, str(d), env)
    a, b = runs["one"], runs["two"]
    assert len(a) == 51 and len(b) == 51
    # the split is time-seeded (yelp.cpp:107), so the two runs see different test sets: same quality band
    m1 = np.mean([row[10] for row in a[-10:]])
    m2 = np.mean([row[10] for row in b[-10:]])
    assert abs(m1 - m2) <= 0.2 * m1, (m1, m2)
    assert b[-1][2] < b[1][2]
