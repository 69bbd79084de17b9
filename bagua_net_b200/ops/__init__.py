"""Hand-written sm_100a ops exposed to Python (thin wrappers over libnccl-net.so):

* ``all_reduce`` / ``all_reduce_oneshot`` — NVLS multimem and NVLink P2P all-reduce
* ``fused_allreduce_sgd``               — gradient all-reduce + optimizer + parameter broadcast
* ``pack_cast``                          — multi-tensor pack with dtype cast into a flat buffer
* ``ConvBiasReLU``                       — conv + fused bias/ReLU/max-pool forward and backward passes
* ``p2p``                                — the NVLink transport's copy / reduce / cast executor
* ``tc_linear``                          — tcgen05/TMEM/TMA linear (+bias+ReLU) and the fused row-parallel GEMM + all-reduce
                                           (opt-in: BNET_TC=1; import it as ``bagua_net_b200.ops.tc_linear``)
"""
from .collectives import all_reduce, all_reduce_oneshot, fused_allreduce_sgd, pack_cast  # noqa: F401
from .fused_nn import ConvBiasReLU  # noqa: F401
from .p2p import P2PExecutor  # noqa: F401
