"""Data-parallel VGG16 training on the fused path: backward overlaps ONE kernel per gradient bucket that all-reduces,
applies SGD to this rank's fp32 shard and broadcasts the new bf16 parameters (bagua_net_b200.parallel.BnetDDP).
The workload is the reference's headline benchmark (reference README.md:52-84), with synthetic ImageNet-shaped data."""
import argparse
import os

import torch

from bagua_net_b200.models import build_model
from bagua_net_b200.parallel import BnetDDP, init_process_group_from_env


def batches(n, batch, device_rank):
    g = torch.Generator().manual_seed(1234 + device_rank)
    for _ in range(n):
        yield (torch.randn(batch, 3, 224, 224, generator=g).bfloat16().pin_memory(),
               torch.randint(0, 1000, (batch,), generator=g).pin_memory())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--model", default="vgg16")
    ap.add_argument("--checkpoint", default="")
    a = ap.parse_args()
    init_process_group_from_env()
    rank = int(os.environ.get("RANK", "0"))
    model = build_model(a.model, fused=True).cuda().bfloat16().to(memory_format=torch.channels_last)
    engine = BnetDDP(model, lr=0.01, momentum=0.9, weight_decay=1e-4)
    engine.enable_cuda_graph(True)            # forward + backward + every bucket kernel replay as one graph
    engine.set_lr(0.01)                       # hyper-parameters live in device memory: the graph follows the schedule
    for step, loss in enumerate(engine.train_from_host(batches(a.steps, a.batch, rank))):
        if step % 10 == 0:
            engine.set_lr(0.01 * 0.5 ** (step // 20))
            if rank == 0:
                print(f"step {step:4d}  loss {float(loss):.4f}")
    if a.checkpoint and rank == 0:
        torch.save({"model": model.state_dict(), "optimizer": engine.optimizer_state_dict()}, a.checkpoint)


if __name__ == "__main__":
    main()
