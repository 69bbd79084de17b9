"""All-reduce of host tensors over the plugin's own connections — no GPU, no NCCL: the deployment the reference was written
for (hosts connected by Ethernet), with the collective done by the transport itself.

    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 4 examples/host_allreduce_over_tcp.py
    BNET_NVL=0 torchrun ... (same)        # force multi-stream TCP even between processes of one host
    BAGUA_NET_IMPLEMENT=TOKIO BAGUA_NET_NSTREAMS=4 ...   # the reference's knobs apply

torch.distributed (gloo) only carries the 128-byte connection handles; the data moves through `TransportMesh`: reduce-scatter
into the slice owners, then all-gather (two network steps for any world size).  Between hosts every piece is a multi-stream
TCP message (the reference's striping); between processes of one host it rides the shared-memory ring / single-copy path."""
import time

import torch
import torch.distributed as dist

from bagua_net_b200.parallel.transport_ring import TransportMesh


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    if world < 2:
        print("needs at least two ranks (torchrun --nproc-per-node 2 ...)")
        return
    mesh = TransportMesh()
    n = 8 << 20
    g = mesh.buffer(n, torch.float32, device="cpu")
    g.fill_(rank + 1)
    mesh.all_reduce(g, g, algo="two-shot")
    assert float(g[0]) == float(g[-1]) == world * (world + 1) / 2
    dist.barrier()
    t0 = time.perf_counter()
    iters = 5
    for _ in range(iters):
        mesh.all_reduce(g, g, algo="two-shot")
    dt = (time.perf_counter() - t0) / iters
    # the same all-reduce through torch.distributed's own CPU backend (gloo, TCP), for scale
    h = torch.full((n,), float(rank + 1))
    dist.all_reduce(h)
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(iters):
        dist.all_reduce(h)
    dt_gloo = (time.perf_counter() - t0) / iters
    if rank == 0:
        bus = n * 4 * 2 * (world - 1) / world / 1e9
        print(f"world {world} over '{mesh.transport}': 32 MiB fp32 all-reduce in {dt * 1e3:.1f} ms ({bus / dt:.2f} GB/s bus bandwidth); "
              f"gloo: {dt_gloo * 1e3:.1f} ms ({bus / dt_gloo:.2f} GB/s)")
    dist.barrier()
    mesh.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
