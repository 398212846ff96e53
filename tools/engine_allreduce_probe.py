"""Isolates the engine's per-minibatch all-reduce cost: tiny user count (compute ~ 0), the
config-B item table (12.8 MB gradient buffer), many minibatches per epoch.  torchrun, 2+ GPUs."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from cdae_b200 import CDAE, CDAEConfig, synth
    from cdae_b200.dist import init_process_group_engine
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    U, I, B = 4096, 50000, 128
    d = synth.make_dataset(U, I, 30.0, seed=1)
    m = CDAE(CDAEConfig(batch_users=B, device=local, loss="CE", num_dim=50, beta=1.0)).reset(
        U, I, d["train_row_ptr"], d["train_col"])
    init_process_group_engine(m)
    m.init_params(1)
    for prof in (False, True):
        m.profile(prof)
        for e in range(3):
            dist.barrier(); torch.cuda.synchronize()
            t = time.perf_counter()
            st = m.train_one_iteration(seed=1, epoch=e)
            w = time.perf_counter() - t
            if rank == 0:
                print("profile=%s epoch %d: %d minibatches wall %.3f ms device %.3f ms -> %.1f us/minibatch"
                      % (prof, e, U // B, w * 1e3, st.device_ms, st.device_ms * 1e3 / (U // B)), flush=True)
        if prof and rank == 0:
            print({k: (round(v[0], 3), v[1]) for k, v in m.profile_get().items() if v[1]}, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
