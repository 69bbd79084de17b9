"""The fused collectives without any framework around them: a symmetric heap shared by all ranks of one NVSwitch box and
hand-written kernels on it (csrc/cuda/coll.cu) — in-switch NVLS reduction, NVLink peer loads/stores, one-shot small
messages — plus the same operations on ordinary CUDA tensors."""
import torch

from bagua_net_b200.parallel import SymmComm, init_process_group_from_env


def main():
    init_process_group_from_env()
    comm = SymmComm(256 << 20)
    g = comm.alloc(16 << 20, torch.bfloat16)              # lives in the heap: kernels reduce it in place
    g.fill_(comm.rank + 1)
    comm.all_reduce(g, algo="auto")                       # NVLS multimem when the box has it, P2P otherwise
    torch.cuda.synchronize()
    expect = comm.world * (comm.world + 1) / 2
    assert float(g[0]) == expect and float(g[-1]) == expect
    t = torch.randn(1000003, device="cuda")               # any tensor: staged through the heap in chunks
    s = comm.all_reduce_tensor(t.clone())
    parts = comm.all_gather_tensor(t[:10])
    shard = comm.reduce_scatter_tensor(torch.ones(comm.world * 8, device="cuda"))
    torch.cuda.synchronize()
    if comm.rank == 0:
        print(f"world {comm.world}  multicast {comm.has_multicast}  all_reduce ok ({expect})  "
              f"gathered {tuple(parts.shape)}  shard {shard.tolist()}  sum[0] {float(s[0]):.3f}")
    comm.close()


if __name__ == "__main__":
    main()
