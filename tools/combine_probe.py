"""Pure cost of the per-minibatch combine step (no user work, no rank skew) per mode and table shape.
torchrun --nproc-per-node N tools/combine_probe.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from cdae_b200 import CDAE, CDAEConfig, _lib

def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    def gather(b):
        box = [None] * world
        dist.all_gather_object(box, b)
        return box
    out = {}
    shapes = (("B", 50_000, 50), ("D", 200_000, 100), ("E", 100_000, 256))
    if os.environ.get("PROBE_SHAPES"):
        shapes = tuple(x for x in shapes if x[0] in os.environ["PROBE_SHAPES"].split(","))
    for name, I, K in shapes:
        for mode in ("nccl", "p2p", "nvls"):
            os.environ.pop("CDAE_B200_DEBUG_NOZERO", None)
            if mode.endswith("-nozero"):
                os.environ["CDAE_B200_DEBUG_NOZERO"] = "1"
                mode_real = mode[:-7]
            else:
                mode_real = mode
            U = 64 * world
            rp = np.arange(U + 1, dtype=np.int64) * 2
            col = np.tile(np.array([0, 1], np.int32), U)
            m = CDAE(CDAEConfig(loss="CE", num_dim=K, beta=1.0, asymmetric=(name == "E"), device=local, batch_users=64 * world)).reset(U, I, rp, col)
            uid = [CDAE.dist_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            m.dist_init(rank, world, uid[0])
            if mode_real == "p2p":
                m.dist_p2p_init(gather)
            elif mode_real == "nvls" and not m.dist_mc_init(rank, world, gather):
                m.close(); continue
            m.init_params(1)
            _lib.check(m._L.cdae_debug_combine(m._h, 5))
            dist.barrier()
            m.profile(True)
            reps = 20
            _lib.check(m._L.cdae_debug_combine(m._h, reps))
            p = m.profile_get()
            us = 1e3 * (p["allreduce"][0] + p["apply"][0]) / reps
            t = torch.tensor([us], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out["%s %s" % (name, mode)] = round(t.item(), 1)
            if mode_real != "nccl":
                import ctypes as C
                ts = (C.c_uint64 * 5)()
                _lib.check(m._L.cdae_debug_combine_times(m._h, ts))
                ph = [round((ts[i + 1] - ts[i]) / 1e3, 1) for i in range(4)]
                allp = gather(ph)
                if rank == 0:
                    out["%s %s phases_us [wait peers, slice work, other blocks, closing barrier] per rank" % (name, mode)] = allp
            m.close()
    if rank == 0:
        print(json.dumps({"world": world, "combine_plus_apply_us_max_over_ranks": out}))
    dist.destroy_process_group()
main()
