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


def make_dataset(U, I, mean_train=30.0, seed=DEFAULT_SEED, test_ratio=0.2, alpha=1.0, item_seed=None):
    """Returns dict(train_row_ptr, train_col, test_row_ptr, test_col, U, I).

    ``item_seed`` (optional) draws the popularity permutation of the items from its own stream, so
    that blocks of users generated with different ``seed`` share ONE item popularity model
    (make_sharded_dataset)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_full = _draw_lengths(rng, U, I, mean_train / (1.0 - test_ratio))
    if item_seed is None:
        perm = rng.permutation(I).astype(np.int64)        # popularity rank -> item id
    else:
        perm = np.random.Generator(np.random.PCG64(item_seed)).permutation(I).astype(np.int64)
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


def owned_users(U, batch_users, rank, world):
    """Global ids of the users rank `rank` trains: its contiguous slice of every global minibatch of
    `batch_users` users — the rule of build_plan (csrc/api.cu) and cdae_b200/dist.py."""
    out = []
    for lo in range(0, U, batch_users):
        n = min(batch_users, U - lo)
        out.append(np.arange(lo + n * rank // world, lo + n * (rank + 1) // world, dtype=np.int64))
    return np.concatenate(out) if out else np.zeros(0, np.int64)


def make_sharded_dataset(U, I, mean_train, batch_users, rank, world, seed=DEFAULT_SEED):
    """The part of a U x I data set that ONE rank of a data-parallel group needs ("each
    GPU holds its CSR shard"): the rows of the users it owns, generated as an independent block
    (seed + rank) over the shared item popularity model (item_seed = seed).  Returned as a GLOBAL
    CSR in which the rows of users owned by other ranks are empty, which is what cdae_create /
    cdae_train_epoch_csr take in a process group (they only ever read the rows a rank trains).
    world == 1 is make_dataset(U, I, mean_train, seed) with the shared item stream."""
    own = owned_users(U, batch_users, rank, world)
    d = make_dataset(len(own), I, mean_train, seed=seed + rank, item_seed=seed)
    out = dict(U=U, I=I, owned=own)
    for part in ("train", "test"):
        lens = np.zeros(U, np.int64)
        lens[own] = np.diff(d[part + "_row_ptr"])
        out[part + "_row_ptr"] = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        out[part + "_col"] = d[part + "_col"]            # owned users ascend with their local index
    return out


def make_blocked_dataset(U, I, mean_train, seed=DEFAULT_SEED, block_users=16384, workers=None):
    """make_dataset for large U: blocks of `block_users` users generated independently (seed + block
    index) over one shared item popularity model and concatenated; blocks run on a thread pool (numpy's
    sort / unique release the GIL).  The cost of make_dataset is dominated by sorting all (user, item)
    keys at once, so this is several times faster at config C / D sizes and gives the same
    distribution (not the same arrays) as one call."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    nb = (U + block_users - 1) // block_users
    sizes = [min(block_users, U - b * block_users) for b in range(nb)]
    workers = workers or min(nb, os.cpu_count() or 1, 32)
    with ThreadPoolExecutor(workers) as ex:
        parts = list(ex.map(lambda b: make_dataset(sizes[b], I, mean_train, seed=seed + b, item_seed=seed), range(nb)))
    out = dict(U=U, I=I)
    for part in ("train", "test"):
        lens = np.concatenate([np.diff(p[part + "_row_ptr"]) for p in parts])
        out[part + "_row_ptr"] = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        out[part + "_col"] = np.concatenate([p[part + "_col"] for p in parts])
    return out
