"""One rank of the torch-facing mesh test on the CPU: torch.distributed (gloo) carries the connection handles, the all-reduce
itself rides the plugin's connections (TCP or shared memory) on host tensors.
usage: tmesh_torch_worker.py   (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT in the environment)"""
import json
import os

import torch
import torch.distributed as dist

from bagua_net_b200.parallel.transport_ring import TransportMesh

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
mesh = TransportMesh()
ok = True
g = mesh.buffer(50000, torch.float32, device="cpu")
for rnd in range(2):
    g.copy_(((torch.arange(50000) * 3 + rank + rnd) % 11 - 5).float())
    dist.barrier()
    mesh.all_reduce(g, g, algo="two-shot")
    want = sum(((torch.arange(50000) * 3 + r + rnd) % 11 - 5).float() for r in range(world))
    ok = ok and bool(torch.equal(g, want))
    dist.barrier()
x, y = mesh.buffers(4097, torch.bfloat16, torch.float32, device="cpu")
x.copy_(((torch.arange(4097) + rank) % 5 - 2).to(torch.bfloat16))
dist.barrier()
mesh.all_reduce(x, y, algo="two-shot")
ok = ok and bool(torch.equal(y, sum(((torch.arange(4097) + r) % 5 - 2).float() for r in range(world))))
try:
    mesh.all_reduce(x, y)                      # one-shot needs device memory
    ok = False
except ValueError:
    pass
print(json.dumps({"ok": ok, "transport": mesh.transport}))
dist.barrier()
mesh.close()
dist.destroy_process_group()
