#!/bin/bash
# 8-GPU session: correctness at world 8, flagship bench (ours / stock NCCL / NCCL over the plugin),
# all-reduce sweeps vs stock NCCL, nccl-tests-style sweep with and without the plugin.
TAG=${1:-s8}; NG=${2:-8}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD
exec > >(tee $OUT/session.log) 2>&1
echo "== session $TAG ngpus=$NG $(date -u)"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv | head -10
step() { local name=$1 tmo=$2; shift 2; echo "---- [$name] $(date -u +%T)"; timeout -k 5 $tmo "$@" > $OUT/$name.log 2>&1; local rc=$?; echo "---- [$name] rc=$rc"; grep -v "Warning\|warn\|^$\|\*\*\*\|OMP_NUM\|return func" $OUT/$name.log | tail -${TAILN:-12} | cut -c1-${CUT:-300}; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
step build 300 make -j16
TAILN=10 step w8_allreduce 200 $TR --master-port 29501 tests/gpu_worker.py allreduce
TAILN=10 step w8_fused_sgd 120 $TR --master-port 29502 tests/gpu_worker.py fused_sgd
CUT=1800 TAILN=2 step bench8 300 $TR --master-port 29541 bench.py --gpus $NG --steps 30 --warmup 5
CUT=1800 TAILN=2 step bench8_nccl 300 $TR --master-port 29542 bench.py --gpus $NG --steps 20 --warmup 5 --comm nccl --no-e2e
TAILN=90 step sweep_all 400 $TR --master-port 29562 bench/allreduce_sweep.py --min-bytes 1K --max-bytes 1G --json $OUT/sweep_all.json
TAILN=40 step sweep_blocks 300 $TR --master-port 29561 bench/allreduce_sweep.py --min-bytes 256M --max-bytes 256M --algos nvls,p2p --blocks 16,32,64,96,148,296 --json $OUT/sweep_blocks.json
TAILN=16 step nccl_perf_stock 200 build/bench/all_reduce_perf -b 8 -e 128M -f 4 -N $NG -d bfloat16
TAILN=16 step nccl_perf_plugin 110 env $(python -m bagua_net_b200.utils.env) build/bench/all_reduce_perf -b 8 -e 128M -f 4 -N $NG -d bfloat16
CUT=1800 TAILN=2 step bench8_resnet50 300 $TR --master-port 29544 bench.py --gpus $NG --steps 20 --warmup 5 --model resnet50 --no-extra
echo "== done $(date -u)"
