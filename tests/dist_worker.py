"""Multi-GPU parity worker: launched by tests/test_gpu_dist.py under torchrun (one process per GPU).

Every rank trains the same problem data-parallel (users of each minibatch sharded over ranks, ONE
NCCL all-reduce of the dense item-side gradients per minibatch) and rank 0
compares the result with the CPU oracle's frozen-batch epoch on the SAME global minibatches:
the set of per-user gradients does not depend on the GPU count, only the fp32 summation order.
Also checks data_loss (sum of per-rank partials) and the per-rank top-N tables.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from cdae_b200 import CDAE, CDAEConfig
    from cdae_b200.dist import owned_users
    from oracle import oracle as orc
    from tests import cases

    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    failures = []
    try:
        _run(rank, world, local, failures)
    except Exception as e:  # noqa: BLE001 - report, then still join the final collective
        import traceback
        print("[rank %d] FAIL exception: %r\n%s" % (rank, e, traceback.format_exc()), flush=True)
        os._exit(3)          # peers are blocked inside NCCL: die loudly, torchrun tears the group down
    flag = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(flag)
    for f in failures:
        print("[rank %d] FAIL %s" % (rank, f), flush=True)
    if rank == 0:
        print("dist_worker: world=%d %s" % (world, "ok" if flag.item() == 0 else "FAILED"), flush=True)
    dist.destroy_process_group()
    return 1 if flag.item() else 0


def _run(rank, world, local, failures):
    import torch.distributed as dist
    from cdae_b200 import CDAE, CDAEConfig
    from cdae_b200.dist import owned_users
    from oracle import oracle as orc
    from tests import cases
    # the last case is full-item-decode training (H12): every rank scores its users against all items
    # on tcgen05 and adds n_rank*lambda*W' to its gradient, the all-reduce makes that n*lambda*W'
    for kw in (dict(loss="CE", beta=1.0, num_dim=50), dict(loss="SQUARE", asymmetric=True, num_dim=20),
               dict(loss="CE", user_factor=False, num_dim=33),
               dict(loss="CE", beta=1.0, asymmetric=True, num_dim=100, full_decode=True),
               # the combine step over NVLink peer memory instead of NCCL: reduce-scatter by peer loads, each
               # rank's slice of the optimiser step (sharded AdaGrad state), all-gather by peer stores
               dict(loss="CE", beta=1.0, num_dim=50, p2p=True),
               dict(loss="SQUARE", asymmetric=True, num_dim=20, p2p=True),
               dict(loss="CE", using_adagrad=False, learn_rate=0.02, num_dim=33, p2p=True),
               dict(loss="CE", beta=1.0, asymmetric=True, num_dim=100, full_decode=True, p2p=True),
               # the same fused step through the NVSwitch multicast engine (skipped where NVLS is unavailable)
               dict(loss="CE", beta=1.0, num_dim=50, nvls=True),
               dict(loss="SQUARE", asymmetric=True, num_dim=20, nvls=True),
               dict(loss="CE", beta=1.0, asymmetric=True, num_dim=100, full_decode=True, nvls=True),
               # every rank holds only the rows of the users it trains (the others are empty in its CSR)
               dict(loss="CE", beta=1.0, num_dim=50, p2p=True, sharded_csr=True),
               dict(loss="CE", beta=1.0, num_dim=50, sharded_csr=True)):
        kw = dict(kw)
        full = kw.pop("full_decode", False)
        use_p2p = kw.pop("p2p", False)
        use_mc = kw.pop("nvls", False)
        sharded = kw.pop("sharded_csr", False)
        cfg = orc.default_config(**kw)
        data = cases.small_dataset(U=403, I=500, mean=12.0, seed=5)
        U, I, K = data["U"], data["I"], cfg["num_dim"]
        rp, col = data["train_row_ptr"], data["train_col"]
        p = cases.random_params(U, I, K, 9, cfg["asymmetric"], cfg["user_factor"])
        B = 96                                            # global minibatch (not a multiple of world)
        my_rp, my_col = rp, col
        if sharded:
            own = np.zeros(U, bool)
            own[owned_users(U, B, rank, world)] = True
            lens = np.where(own, np.diff(rp), 0)
            my_rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
            my_col = np.concatenate([col[rp[u]:rp[u + 1]] for u in range(U) if own[u]]).astype(np.int32)
        m = CDAE(CDAEConfig(batch_users=B, device=local, full_decode=full, **cfg)).reset(U, I, my_rp, my_col)
        uid = [CDAE.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        m.dist_init(rank, world, uid[0])
        def gather(b):
            box = [None] * world
            dist.all_gather_object(box, b)
            return box
        if use_mc:
            if not m.dist_mc_init(rank, world, gather):
                if rank == 0:
                    print("dist_worker: NVLS unavailable here (%s) - case skipped" % kw, flush=True)
                m.close()
                continue
        if use_p2p:
            m.dist_p2p_init(gather)
        m.set_params(p)
        steps = 0
        for epoch in range(3):
            # the third epoch takes the training set from host memory (cdae_train_epoch_csr): a rank
            # uploads only the rows of the users it trains
            st = m.train_one_iteration(seed=123, epoch=epoch, csr=(my_rp, my_col) if epoch == 2 else None)
            steps += st.user_steps
        mine = owned_users(U, B, rank, world)
        if steps != 3 * len(mine):
            failures.append("rank %d trained %d user steps, owns %d users" % (rank, steps, len(mine)))
        # one more frozen minibatch with EXPLICIT masks / negatives (cdae_train_users is collective in a group:
        # every rank passes the same lists and trains its slice of them)
        ex_users = np.arange(7, 7 + 81)
        ex = cases.draw_step_inputs(data, cfg["num_neg"], cfg["corruption_ratio"], np.random.default_rng(77), ex_users)
        if not sharded:
            ex_keep = np.concatenate([ex[u][0] for u in ex_users]).astype(np.uint8)
            ex_negs = None if full else np.concatenate([ex[u][1] for u in ex_users]).astype(np.int32)
            m.train_users(ex_users, ex_keep, ex_negs)
        got = {k: m.get_param(k) for k in ("W", "V", "Wu", "b", "b_prime", "W_ag", "V_ag", "Wu_ag", "b_ag", "b_prime_ag")}
        pen = m.penalty_loss()
        rows = np.array([0, 5, 17, 200, U - 1])
        got_rows = m._get_rows("Wu", rows) if cfg["user_factor"] else None
        loss = m.data_loss(seed=7) if not sharded else None
        ids, _ = m.recommend_all(10)
        # checkpoint round trip: collective save (rank 0 writes), every rank loads
        path = [os.path.join("/tmp", "cdae_ckpt_%d.bin" % os.getpid()) if rank == 0 else None]
        dist.broadcast_object_list(path, src=0)
        m.save(path[0])
        dist.barrier()
        m.set_params({"b": np.zeros(K)})
        m.load(path[0])
        back = {k: m.get_param(k) for k in ("W", "Wu", "b", "W_ag")}
        for k, v in back.items():
            if v.size and not np.array_equal(v, got[k]):
                failures.append("%s rank %d: %s differs after save/load" % (kw, rank, k))
        dist.barrier()
        if rank == 0:
            os.unlink(path[0])
        if rank == 0:
            o = orc.Oracle(cfg, U, I, rp, col)
            o.set_params(p)
            for epoch in range(3):
                if full:
                    o.train_epoch_full(123, epoch, B, rounding=1)     # restates the bf16 operand rounding
                else:
                    o.train_epoch(123, epoch, batch_users=B)
            if not sharded:
                ins = [col[rp[u]:rp[u + 1]][ex[u][0]] for u in ex_users]
                if full:
                    o.step_frozen_full(ex_users, ins, rounding=1)
                else:
                    o.step_frozen(ex_users, ins, [ex[u][1] for u in ex_users])
            if not abs(pen - o.penalty_loss()) <= 1e-4 * o.penalty_loss() * (30 if full else 1):
                failures.append("%s penalty_loss %.6f vs %.6f" % (kw, pen, o.penalty_loss()))
            if got_rows is not None and not np.allclose(got_rows, o.param("Wu")[rows], rtol=3e-3 if full else 2e-4, atol=2e-5):
                failures.append("%s get_param_rows(Wu) differs" % (kw,))
            for k, v in got.items():
                ref = o.param(k)
                if ref.size == 0 or v.size == 0:
                    continue
                err = np.abs(v - ref).max() / max(1e-12, np.abs(ref).max())
                # full decode: three epochs of bf16 rounding flips (tests/test_gpu_fulldec.py) on top of fp32 order
                # (full-decode accumulators: sums of squared bf16-operand sums, 8e-3 as in tests/test_gpu_fulldec.py)
                if not err <= ((8e-3 if k.endswith("_ag") else 3e-3) if full else 2e-4):
                    failures.append("%s%s%s%s %s: max err %.3g" % (kw, " p2p" if use_p2p else "", " nvls" if use_mc else "", " sharded-csr" if sharded else "", k, err))
            keep = np.concatenate([o.sample_keep(7, 0x80000000, u) for u in range(U)])
            ref_loss = o.data_loss(keep)
            if loss is not None and not abs(loss - ref_loss) <= (2e-3 if full else 2e-4) * abs(ref_loss):
                failures.append("%s data_loss %.6f vs %.6f" % (kw, loss, ref_loss))
        # every rank: its own users' lists match the oracle evaluated on ITS (identical) parameters
        o2 = orc.Oracle(cfg, U, I, my_rp, my_col)
        o2.set_params({k: v for k, v in m.get_params().items() if v.size})
        for u in mine[:: max(1, len(mine) // 40)]:
            if ids[u].tolist() != o2.recommend(int(u), 10)[0].tolist():
                failures.append("%s rank %d user %d top-10 differs" % (kw, rank, u))
                break
        m.close()


if __name__ == "__main__":
    sys.exit(main())
