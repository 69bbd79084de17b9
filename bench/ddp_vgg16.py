"""vgg16 synthetic data-parallel training benchmark (the reference quotes VGG16: reference README.md:52-84).
Same flags and JSON line as the top-level bench.py; this wrapper only pins --model vgg16.

    python bench/ddp_vgg16.py --gpus 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 bench/ddp_vgg16.py --gpus 8
    ... --comm nccl | nccl-plugin   for torch DDP over stock NCCL / over NCCL forced through the plugin
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv[1:1] = ["--model", "vgg16"]

import bench as _bench  # noqa: E402  (the repo-root bench.py)

if __name__ == "__main__":
    sys.exit(_bench.main())
