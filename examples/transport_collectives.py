"""All-reduces that ride the plugin's own connections (the transport NCCL uses, without NCCL): the reduction is done by the
SENDING GPU's kernel while it moves the data over NVLink, so no reduce kernel and no staging buffer exist.

  ring       2(n-1) steps, every hop a fused isend (`isend_op(OP_RED_ADD_*)`): bandwidth-optimal, bit-reproducible
  compressed the same ring with bf16 / fp8 on the wire: the quantisation rides the isend, NVLink carries 2-4x fewer bytes
  one-shot   a full mesh, ONE step: every peer's kernel accumulates into every output — latency-optimal up to a few MiB
  two-shot   the same mesh, TWO steps for any world size: reduce-scatter into the slice owners (fused isends), all-gather by
             copy; 2(n-1)/n x size on the wire, in place — the shape an NVSwitch wants, and what the CollNet table gives NCCL

    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 4 examples/transport_collectives.py
"""
import time

import torch
import torch.distributed as dist

from bagua_net_b200.parallel import init_process_group_from_env
from bagua_net_b200.parallel.transport_ring import TransportMesh, TransportRing


def timed(fn, iters=5):
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t0) / iters


def main():
    init_process_group_from_env()
    rank, world = dist.get_rank(), dist.get_world_size()
    if world < 2:
        print("needs at least two ranks (torchrun --nproc-per-node 2 ...)")
        return
    n = 16 << 20
    ring = TransportRing()
    g = ring.buffer(n, torch.float32)
    g.fill_(rank + 1)
    ring.all_reduce(g)
    assert float(g[0]) == world * (world + 1) / 2
    t_ring = timed(lambda: ring.all_reduce(g))
    g.fill_((rank + 1) * 0.25)
    ring.all_reduce_compressed(g, wire="e4m3", scale=4.0)                 # partial sums stay exact in e4m3 here
    assert float(g[-1]) == world * (world + 1) / 8
    t_comp = timed(lambda: ring.all_reduce_compressed(g, wire="e4m3", scale=1.0 / 64))
    mesh = TransportMesh()
    x, y = mesh.buffers(1 << 18, torch.bfloat16, torch.float32)           # bf16 gradients, fp32 sums
    x.fill_(rank + 1)
    mesh.all_reduce(x, y)
    assert float(y[0]) == world * (world + 1) / 2
    t_mesh = timed(lambda: mesh.all_reduce(x, y), iters=20)
    h = mesh.buffer(n, torch.float32)                                     # one registered buffer: in place
    h.fill_(rank + 1)
    mesh.all_reduce(h, h, algo="two-shot")
    assert float(h[0]) == world * (world + 1) / 2 and float(h[-1]) == world * (world + 1) / 2
    t_two = timed(lambda: mesh.all_reduce(h, h, algo="two-shot"))
    if rank == 0:
        bus = 2 * (world - 1) / world
        print(f"world {world} over '{ring.transport}': ring 64 MiB fp32 {n * 4 * bus / t_ring / 1e9:.0f} GB/s busbw | "
              f"e4m3 on the wire {n * 4 * bus / t_comp / 1e9:.0f} GB/s (of fp32 payload) | one-shot 512 KiB bf16->fp32 {t_mesh * 1e6:.0f} us | "
              f"two-shot 64 MiB fp32 in place {n * 4 * bus / t_two / 1e9:.0f} GB/s busbw")
    ring.close()
    mesh.close()


if __name__ == "__main__":
    main()
