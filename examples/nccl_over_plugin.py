"""Plain torch.distributed over NCCL — with NCCL's network traffic carried by this plugin, which is all the reference
does (reference README.md:32-45).  Run under the environment printed by `python -m bagua_net_b200.utils.env`
(LD_LIBRARY_PATH, NCCL_NET_PLUGIN=bnet; inside one box also NCCL_P2P_DISABLE=1 NCCL_SHM_DISABLE=1 so that NCCL has to
use a network at all).  BNET_METRICS_FILE=/tmp/bnet.prom (or BAGUA_NET_PROMETHEUS_ADDRESS, like the reference) exposes the
plugin's isend / irecv byte and latency counters; BNET_TRACE_FILE=/tmp/bnet.trace.json its spans.  With NCCL_DEBUG=INFO look for `NET/Plugin: Loaded net plugin BNet (v8)` and `Using network BNet`."""
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    dist.init_process_group("nccl", rank=rank, world_size=world)
    for nbytes in (1 << 12, 1 << 20, 64 << 20):
        x = torch.full((nbytes // 4,), float(rank + 1), device="cuda")
        dist.all_reduce(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            dist.all_reduce(x)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        if rank == 0:
            print(f"all_reduce {nbytes >> 10:8d} KiB  {dt * 1e6:9.1f} us  busbw {2 * (world - 1) / world * nbytes / dt / 1e9:7.2f} GB/s")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
