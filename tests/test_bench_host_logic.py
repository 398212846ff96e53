"""Host-side logic added in round 2 that needs no GPU: the sharded data sets of the benchmark (a rank's CSR holds
only the rows of the users it owns — the ownership rule of build_plan / cdae_b200.dist), the blocked generator, and
the `config` object both bench arms share."""
import numpy as np

from cdae_b200 import synth
from cdae_b200.dist import owned_users as dist_owned


def test_sharded_dataset_rows_follow_the_ownership_rule():
    U, I, B, world = 5000, 700, 1024, 4
    seen = np.zeros(U, int)
    nnz = 0
    for rank in range(world):
        d = synth.make_sharded_dataset(U, I, 12.0, B, rank, world, seed=3)
        own = synth.owned_users(U, B, rank, world)
        assert np.array_equal(own, np.asarray(dist_owned(U, B, rank, world)))       # same rule as the engine's mirror
        lens = np.diff(d["train_row_ptr"])
        mask = np.zeros(U, bool)
        mask[own] = True
        assert (lens[mask] > 0).all() and (lens[~mask] == 0).all()
        assert d["train_row_ptr"][-1] == len(d["train_col"])
        for u in own[:50]:
            row = d["train_col"][d["train_row_ptr"][u]:d["train_row_ptr"][u + 1]]
            assert (np.diff(row) > 0).all() and row.min() >= 0 and row.max() < I     # ascending, in range
        seen[own] += 1
        nnz += len(d["train_col"])
    assert (seen == 1).all()                                                          # every user owned exactly once
    assert abs(nnz / U - 12.0) < 2.0


def test_blocked_dataset_is_a_valid_csr_with_the_requested_profile():
    d = synth.make_blocked_dataset(3000, 500, 15.0, seed=5, block_users=700, workers=2)
    rp, col = d["train_row_ptr"], d["train_col"]
    assert len(rp) == 3001 and rp[-1] == len(col)
    lens = np.diff(rp)
    assert lens.min() >= 1 and abs(lens.mean() - 15.0) < 2.5
    for u in range(0, 3000, 97):
        assert (np.diff(col[rp[u]:rp[u + 1]]) > 0).all()
    # blocks share one item popularity model: the most popular item overall is popular in every block
    top = np.bincount(col, minlength=500).argmax()
    for b in range(0, 3000, 700):
        blk = col[rp[b]:rp[min(b + 700, 3000)]]
        assert np.bincount(blk, minlength=500)[top] > 0.25 * np.bincount(blk, minlength=500).max()


def test_both_bench_arms_describe_one_workload():
    import bench
    for name in bench.CONFIGS:
        a = bench.config_dict(name, 1, 0)
        b = bench.config_dict(name, 1, 123456)
        assert a == b and a["name"] == name and "workload" in a                      # independent of arm-specific knobs
        assert bench.config_dict(name, 8, 0)["users"] == 8 * bench.CONFIGS[name]["users_per_gpu"]
    k, ld = bench.decode_kernel_name(50)
    assert k.startswith("decode_kernel<8,2") and ld == 64
    assert bench.decode_kernel_name(100)[1] == 128 and bench.decode_kernel_name(200)[1] == 256
