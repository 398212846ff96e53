"""Deterministic synthetic implicit-feedback data in the shape of the reference's input.

The reference ships no dataset (apps/yelp reads ./yelp_10core.txt, yelp.cpp:23, which is not
in the repo), so benchmarks and tests use this generator:

* items per user  n_u = clamp(round(LogNormal(mu = ln(mean) - 0.32, sigma = 0.8)), 10,
  min(I/4, 2000))  — a "10-core"-like heavy-tailed profile;
* items drawn without replacement from Zipf(alpha = 1) over a seeded permutation of [0, I);
* per-user split by the reference's rule (data-inl.hpp:252): floor(0.2 * n_u) items to
  test, the rest to train.  ``mean_train`` is the mean TRAIN row length, i.e. the full
  profile is drawn with mean ``mean_train / 0.8``.

Output is CSR (int64 row_ptr, int32 col, ascending inside each row) — the layout
cdae_create() takes.  numpy's PCG64 stream is platform independent, so the same seed gives
the same arrays everywhere (default seed 20141119 = the reference's, yelp.cpp:29).
"""
import numpy as np

DEFAULT_SEED = 20141119


def _draw_lengths(rng, U, I, mean):
    mu = np.log(mean) - 0.32
    n = np.rint(rng.lognormal(mu, 0.8, size=U)).astype(np.int64)
    hi = max(10, min(I // 4, 2000))
    return np.clip(n, min(10, hi), hi)


def make_dataset(U, I, mean_train=30.0, seed=DEFAULT_SEED, test_ratio=0.2, alpha=1.0):
    """Returns dict(train_row_ptr, train_col, test_row_ptr, test_col, U, I)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_full = _draw_lengths(rng, U, I, mean_train / (1.0 - test_ratio))
    perm = rng.permutation(I).astype(np.int64)            # popularity rank -> item id
    w = 1.0 / np.power(np.arange(1, I + 1, dtype=np.float64), alpha)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]

    # Rejection rounds: oversample, de-duplicate (user, item) pairs, retry only the users
    # still short (heavy users saturate the Zipf head and need several rounds).
    done = []                                              # finished users' keys
    keys = np.zeros(0, np.int64)                           # unique user * I + item, unfinished users
    todo = np.arange(U, dtype=np.int64)
    need = n_full.copy()
    for rnd in range(200):
        if todo.size == 0:
            break
        cnt = (need[todo] * (1.3 + 0.7 * rnd)).astype(np.int64) + 4
        users = np.repeat(todo, cnt)
        ranks = np.searchsorted(cdf, rng.random(users.size), side="right")
        ranks = np.minimum(ranks, I - 1)
        keys = np.unique(np.concatenate([keys, users * I + perm[ranks]]))
        have = np.bincount(keys // I, minlength=U)         # unfinished users only
        need[todo] = np.maximum(n_full[todo] - have[todo], 0)
        fin = need[keys // I] == 0
        done.append(keys[fin])
        keys = keys[~fin]
        todo = todo[need[todo] > 0]
    keys = np.sort(np.concatenate(done + [keys]))
    users = keys // I
    items = (keys - users * I).astype(np.int32)
    have = np.bincount(users, minlength=U)
    start = np.concatenate([[0], np.cumsum(have)])
    # keep the first n_full[u] (ids are a random permutation of popularity, so "first by id"
    # does not favour popular items)
    pos = np.arange(keys.size) - start[users]
    keep = pos < n_full[users]
    users, items = users[keep], items[keep]
    n_u = np.bincount(users, minlength=U)
    start = np.concatenate([[0], np.cumsum(n_u)])

    # split: a per-user random subset of floor(test_ratio * n_u) goes to test
    r = rng.random(users.size)
    order = np.lexsort((r, users))                         # random order inside each user
    rank_in_user = np.empty(users.size, np.int64)
    rank_in_user[order] = np.arange(users.size) - start[users[order]]
    n_test = np.floor(n_u * test_ratio).astype(np.int64)
    is_test = rank_in_user < n_test[users]

    def csr(mask):
        u, it = users[mask], items[mask]
        rp = np.concatenate([[0], np.cumsum(np.bincount(u, minlength=U))]).astype(np.int64)
        return rp, np.ascontiguousarray(it, np.int32)      # still ascending inside each row

    trp, tcol = csr(~is_test)
    erp, ecol = csr(is_test)
    return dict(U=U, I=I, train_row_ptr=trp, train_col=tcol, test_row_ptr=erp, test_col=ecol)


def make_params(U, I, K, seed=DEFAULT_SEED, asymmetric=False, user_factor=True):
    """U[-1,1] * 4*sqrt(6/(I+K)) draws (the reference's init scale, cdae.hpp:112-113) as
    fp32-representable doubles; accumulators are left at their 1e-4 defaults."""
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    s = 4.0 * np.sqrt(6.0 / (I + K))
    p = {"W": (rng.uniform(-1, 1, (I, K)) * s).astype(np.float32).astype(np.float64)}
    if asymmetric:
        p["V"] = (rng.uniform(-1, 1, (I, K)) * s).astype(np.float32).astype(np.float64)
    if user_factor:
        p["Wu"] = (rng.uniform(-1, 1, (U, K)) * s).astype(np.float32).astype(np.float64)
    return p
