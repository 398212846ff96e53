"""Small seeded CDAE problems shared by the oracle / reference / GPU parity tests."""
import numpy as np

from cdae_b200 import synth


def small_dataset(U=64, I=257, mean=12.0, seed=7):
    """Synthetic train/test CSR where every item occurs in train and item ids are numbered by
    first appearance in user-major order — the numbering the reference's loader assigns
    (instance-inl.hpp:22-37), so ids agree between the reference, the oracle and the GPU."""
    d = synth.make_dataset(U, I, mean_train=mean, seed=seed)
    trp, tcol = d["train_row_ptr"], d["train_col"].astype(np.int64)
    erp, ecol = d["test_row_ptr"], d["test_col"].astype(np.int64)
    _, first = np.unique(tcol, return_index=True)
    order = tcol[np.sort(first)]
    new_id = np.full(I, -1, np.int64)
    new_id[order] = np.arange(len(order))
    I2 = len(order)

    def relabel(rp, col, drop_unknown):
        rows = []
        for u in range(U):
            r = new_id[col[rp[u]:rp[u + 1]]]
            if drop_unknown:
                r = r[r >= 0]
            rows.append(np.sort(r))
        rp2 = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
        return rp2, np.concatenate(rows).astype(np.int32)

    trp2, tcol2 = relabel(trp, tcol, False)
    erp2, ecol2 = relabel(erp, ecol, True)
    return dict(U=U, I=I2, train_row_ptr=trp2, train_col=tcol2, test_row_ptr=erp2, test_col=ecol2)


def random_params(U, I, K, seed, asymmetric, user_factor, linear_function=False, scale=None,
                  warm=True):
    """fp32-representable parameters.  ``warm`` also randomises biases and AdaGrad state so
    every term of the update rules is exercised (a fresh model has b = b' = 0)."""
    rng = np.random.default_rng(seed)
    s = scale if scale is not None else 4.0 * np.sqrt(6.0 / (I + K))

    def f32(a):
        return a.astype(np.float32).astype(np.float64)

    p = {"W": f32(rng.uniform(-1, 1, (I, K)) * s)}
    if asymmetric:
        p["V"] = f32(rng.uniform(-1, 1, (I, K)) * s)
    if user_factor:
        p["Wu"] = f32(rng.uniform(-1, 1, (U, K)) * s)
    if linear_function:
        p["Uu"] = f32(1.0 + rng.uniform(-0.3, 0.3, (U, K)))
    if warm:
        p["b"] = f32(rng.uniform(-0.2, 0.2, K))
        p["b_prime"] = f32(rng.uniform(-0.2, 0.2, I))
        p["W_ag"] = f32(1e-4 + rng.uniform(0, 0.5, (I, K)))
        p["b_ag"] = f32(1e-4 + rng.uniform(0, 0.5, K))
        p["b_prime_ag"] = f32(1e-4 + rng.uniform(0, 0.5, I))
        if asymmetric:
            p["V_ag"] = f32(1e-4 + rng.uniform(0, 0.5, (I, K)))
        if user_factor:
            p["Wu_ag"] = f32(1e-4 + rng.uniform(0, 0.5, (U, K)))
        if linear_function:
            p["Uu_ag"] = f32(1e-4 + rng.uniform(0, 0.5, (U, K)))
    return p


def draw_step_inputs(data, num_neg, q, rng, users=None, unique_negs=False):
    """Explicit corruption masks and negatives: keep ~ Bernoulli(1-q) per train slot; negatives
    uniform over the user's non-positives, with replacement (without if unique_negs)."""
    rp, col = data["train_row_ptr"], data["train_col"]
    users = range(data["U"]) if users is None else users
    out = {}
    for u in users:
        row = col[rp[u]:rp[u + 1]].astype(np.int64)
        keep = rng.random(len(row)) > q
        cand = np.setdiff1d(np.arange(data["I"]), row)
        n = len(row) * num_neg
        negs = rng.choice(cand, size=n, replace=not unique_negs) if n else np.zeros(0, np.int64)
        out[u] = (keep, negs.astype(np.int64))
    return out


CONFIG_GRID = [
    dict(),                                                     # struct defaults except loss
    dict(loss="SQUARE"),
    dict(asymmetric=True),
    dict(asymmetric=True, loss="SQUARE", using_adagrad=False),
    dict(using_adagrad=False),
    dict(beta=1.0, corruption_ratio=0.2),
    dict(tanh=True),
    dict(linear=True, loss="SQUARE"),
    dict(user_factor=False),
    dict(scaled=False, corruption_ratio=0.8),
    dict(linear_function=True),
    dict(linear_function=True, asymmetric=True, tanh=True, beta=1.0),
    dict(corruption_ratio=0.0, scaled=False, beta=1.0, loss="SQUARE"),   # apps/yelp defaults
    dict(num_neg=1, num_dim=50),
    dict(num_neg=0),
    dict(loss="LOG"), dict(loss="HINGE"), dict(loss="SQUARED_HINGE"), dict(loss="LOGM"),
    dict(corruption_ratio=1.0, beta=1.0),       # q = 1 with scaled: 1/(1-q) = inf is never multiplied (cdae.hpp:366,377-380)
    dict(corruption_ratio=1.0, scaled=False, asymmetric=True),
]
