"""Times an all-reduce of the engine's gradient-buffer size through torch.distributed (NCCL) —
run under torchrun; prints bus bandwidth per size.  NCCL_DEBUG=INFO shows the transport."""
import os
import time

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for mb in (1, 4, 12.8, 64, 256):
        n = int(mb * (1 << 20) / 4)
        x = torch.ones(n, device="cuda")
        for _ in range(5):
            dist.all_reduce(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dist.all_reduce(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        if rank == 0:
            bus = 2 * (world - 1) / world * n * 4 / (ms * 1e-3) / 1e9
            print("allreduce %.1f MB: %.3f ms, busbw %.1f GB/s" % (mb, ms, bus), flush=True)
    if rank == 0:
        print("p2p access 0->1:", torch.cuda.can_device_access_peer(0, 1) if torch.cuda.device_count() > 1 else None)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
