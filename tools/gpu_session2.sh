#!/bin/bash
# Session 2: re-run tests after fixes, diagnose NCCL-over-plugin layer by layer (short timeouts),
# tune the NVLS kernels, stock-NCCL comparison sweep.   usage: tools/gpu_session2.sh <tag> <ngpus>
TAG=${1:-s2}; NG=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD
exec > >(tee $OUT/session.log) 2>&1
echo "== session $TAG ngpus=$NG $(date -u)"
step() { local name=$1 tmo=$2; shift 2; echo "---- [$name] $(date -u +%T)"; timeout -k 5 $tmo "$@" > $OUT/$name.log 2>&1; local rc=$?; echo "---- [$name] rc=$rc"; tail -${TAILN:-12} $OUT/$name.log | cut -c1-400; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
step build 600 make -j16
TAILN=25 step gpu_tests_1 600 python -m pytest tests/test_gpu.py -q -m "gpu and not multigpu" -p no:cacheprovider
TAILN=40 step gpu_tests_multi 600 python -m pytest tests/test_gpu.py -q -m "multigpu" -p no:cacheprovider -k "not nccl_loads" -s
# ---- NCCL in the loop, one layer at a time
PENV=$(python -m bagua_net_b200.utils.env --debug)
nccl_case() { local name=$1; shift; TAILN=30 step $name 75 env $PENV BNET_LOG_LEVEL=INFO BNET_WATCHDOG_MS=8000 "$@" $TR --master-port 29551 tests/gpu_worker.py nccl_allreduce; grep -E "Using network|Loaded net plugin|GDR|via NET|BNet|bnet" $OUT/$name.log | head -20 | cut -c1-300; }
nccl_case nccl_tcp_host BNET_NVL=0 BNET_GDR=0
nccl_case nccl_shm_host BNET_NVL=1 BNET_GDR=0
nccl_case nccl_nvl_gdr_oneshot BNET_NVL=1 BNET_PERSISTENT=0
nccl_case nccl_nvl_gdr_persistent BNET_NVL=1
# ---- flagship bench
step bench1 300 python bench.py --gpus 1 --steps 20 --warmup 5
step bench2 400 $TR --master-port 29541 bench.py --gpus $NG --steps 20 --warmup 5
# ---- tuning: CTA count for the NVLS all-reduce, engine choice for the transport kernels
TAILN=60 step sweep_blocks 400 $TR --master-port 29561 bench/allreduce_sweep.py --min 16M --max 1G --algos nvls,p2p --blocks 32,64,96,148,200,296 --json $OUT/sweep_blocks.json
TAILN=60 step sweep_all 400 $TR --master-port 29562 bench/allreduce_sweep.py --min 1K --max 1G --json $OUT/sweep_all.json
TAILN=20 step p2p_ldst 200 $TR --master-port 29563 bench/p2p_bw.py
TAILN=20 step p2p_tma 200 env BNET_COPY_ENGINE=tma $TR --master-port 29564 bench/p2p_bw.py
TAILN=20 step p2p_8x4 200 env BNET_NCLUSTERS=8 BNET_CLUSTER_SIZE=4 $TR --master-port 29565 bench/p2p_bw.py
TAILN=40 step nccl_perf_stock 300 build/bench/all_reduce_perf -b 8 -e 128M -f 4 -N $NG -d bfloat16
echo "== done $(date -u)"
