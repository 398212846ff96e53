"""world_size = 2 on CPU (gloo): the host-side logic of the data-parallel path.

No CUDA here — the kernels are covered by tests/test_gpu_dist.py on real GPUs.  What runs on the
CPU is everything around them: the shard rule (cdae_b200.dist, mirrors build_plan in csrc/api.cu),
the reduction algebra (per-rank frozen gradients, summed with a real all-reduce, applied once ==
the single-process frozen step), ownership of the user-private rows, and the rendezvous that
hands rank 0's 128-byte NCCL id to every rank.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_rule_partitions_every_minibatch():
    from cdae_b200.dist import minibatch_slices, owned_users
    for U, B, world in [(403, 96, 2), (1000, 128, 8), (7, 8192, 4), (100, 1, 3), (64, 64, 2)]:
        per_rank = [minibatch_slices(U, B, r, world) for r in range(world)]
        n_mb = (U + B - 1) // B
        assert all(len(s) == n_mb for s in per_rank)          # same number of all-reduces everywhere
        for mb in range(n_mb):
            lo, hi = mb * B, min(U, (mb + 1) * B)
            assert per_rank[0][mb][0] == lo and per_rank[-1][mb][1] == hi
            for r in range(world - 1):
                assert per_rank[r][mb][1] == per_rank[r + 1][mb][0]   # contiguous, disjoint
            sizes = [b - a for a, b in (per_rank[r][mb] for r in range(world))]
            assert max(sizes) - min(sizes) <= 1                # balanced
        allu = np.sort(np.concatenate([owned_users(U, B, r, world) for r in range(world)]))
        assert allu.tolist() == list(range(U))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from cdae_b200.dist import minibatch_slices, owned_users
    from oracle import oracle as orc
    from tests import cases
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rendezvous: a 128-byte id made on rank 0 reaches every rank unchanged
        box = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        assert box[0] == bytes(range(128))

        out = {}
        # the last two cases are full-item-decode training (H12): a rank's gradient carries
        # n_rank * lambda * W' for EVERY item, and the sum over ranks must be the minibatch's n * lambda * W'
        for kw in (dict(loss="CE", beta=1.0), dict(loss="SQUARE", asymmetric=True, using_adagrad=False),
                   dict(loss="CE", linear_function=True),
                   dict(loss="CE", beta=1.0, asymmetric=True, full=True), dict(loss="SQUARE", beta=1.0, full=True)):
            kw = dict(kw)
            full = kw.pop("full", False)
            cfg = orc.default_config(**kw)
            data = cases.small_dataset(U=150, I=300, mean=12.0, seed=11)
            U, I, K = data["U"], data["I"], cfg["num_dim"]
            rp, col = data["train_row_ptr"], data["train_col"]
            p = cases.random_params(U, I, K, 4, cfg["asymmetric"], cfg["user_factor"],
                                    cfg["linear_function"])
            B, seed, cnum = 64, 99, cfg["num_corruptions"]
            o = orc.Oracle(cfg, U, I, rp, col)
            o.set_params(p)
            for epoch in range(2):
                for (a, b) in minibatch_slices(U, B, rank, world):
                    users = np.arange(a, b)
                    dense = np.zeros(o.dense_grad_size())
                    keeps = [o.sample_keep(seed, epoch * cnum, u) for u in users]
                    ins = [col[rp[u]:rp[u + 1]][k.astype(bool)] for u, k in zip(users, keeps)]
                    if full:
                        o.shard_gradients_full(users, ins, dense)
                    else:
                        negs = [o.sample_negatives(seed, epoch * cnum, u) for u in users]
                        o.shard_gradients(users, ins, negs, dense)
                    t = torch.from_numpy(dense)
                    dist.all_reduce(t)                         # the ONE collective per minibatch
                    o.apply_dense(dense)
            # user-private rows: keep owned rows, zero the rest, sum over ranks (get_param's rule)
            mine = owned_users(U, B, rank, world)
            res = {}
            for k in ("W", "V", "b", "b_prime", "W_ag", "b_prime_ag"):
                res[k] = o.param(k).copy()
            for k in ("Wu", "Wu_ag", "Uu", "Uu_ag"):
                a = o.param(k)
                if a.size == 0:
                    res[k] = a.copy()
                    continue
                own = np.zeros_like(a)
                own[mine] = a[mine]
                t = torch.from_numpy(own)
                dist.all_reduce(t)
                res[k] = own
            # replicated blocks must be bit-identical on every rank
            for k in ("W", "b", "b_prime"):
                t = torch.from_numpy(res[k].copy())
                dist.broadcast(t, src=0)
                assert np.array_equal(t.numpy(), res[k]), "replica drift in " + k
            if rank == 0:
                single = orc.Oracle(cfg, U, I, rp, col)
                single.set_params(p)
                for epoch in range(2):
                    if full:
                        single.train_epoch_full(seed, epoch, B)
                    else:
                        single.train_epoch(seed, epoch, batch_users=B)
                for k, v in res.items():
                    ref = single.param(k)
                    if ref.size:
                        err = float(np.abs(v - ref).max() / max(1.0, np.abs(ref).max()))
                        out["%s%s/%s" % (sorted(kw.items()), " full" if full else "", k)] = err
        q.put((rank, "ok", out))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "error: %r\n%s" % (e, traceback.format_exc()), {}))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_reduction_equals_single_process(oracle_built):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, _ in results:
        assert status == "ok", "rank %d: %s" % (rank, status)
    errs = [d for r, _, d in results if r == 0][0]
    assert errs, "rank 0 compared nothing"
    for k, e in errs.items():
        assert e <= 1e-9, (k, e)           # fp64 on both sides: only the summation order differs
