#!/bin/bash
# Round-2 session 3 (2 GPUs): batched per-message launches, the DDP arms one by one (with stack dumps), the full bench,
# tcgen05 conv vs cuDNN per layer shape, the multi-GPU tests that have not run yet.
TAG=${1:-r2s3}; NG=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD
exec > >(tee $OUT/session.log) 2>&1
echo "== session $TAG ngpus=$NG $(date -u)"
make -j16 >/dev/null 2>&1; make bench >/dev/null 2>&1
BASE="$(python -m bagua_net_b200.utils.env) BNET_WATCHDOG_MS=5000 NCCL_DEBUG=WARN"
ARP="build/bench/all_reduce_perf -N $NG -d bfloat16"
run() { local name=$1 tmo=$2 pargs=$3; shift 3; echo "---- [$name] $(date -u +%T) $*"; timeout -k 5 $tmo env $BASE "$@" $ARP $pargs > $OUT/$name.log 2>&1; local rc=$?; echo "---- [$name] rc=$rc"; grep -v "^$" $OUT/$name.log | tail -${TAILN:-16} | cut -c1-330; return $rc; }
SWEEP="-b 8 -e 128M -f 4 -n 10 -w 3"
run batched 60 "$SWEEP"
run unbatched 60 "$SWEEP" BNET_MSG_BATCH=0
TAILN=8 run batched_c8 45 "-b 1M -e 128M -f 8 -n 10 -w 3" NCCL_MIN_NCHANNELS=8
TAILN=8 run batched_c32 45 "-b 1M -e 128M -f 8 -n 10 -w 3" NCCL_MIN_NCHANNELS=32
TAILN=16 run ll_allowed 60 "$SWEEP" NCCL_PROTO=LL,Simple
step() { local name=$1 tmo=$2; shift 2; echo "---- [$name] $(date -u +%T)"; timeout -k 5 $tmo "$@" > $OUT/$name.log 2>&1; local rc=$?; echo "---- [$name] rc=$rc"; grep -v "Warning\|warn\|^$\|\*\*\*\|OMP_NUM\|return func" $OUT/$name.log | tail -${TAILN:-8} | cut -c1-${CUT:-2500}; return $rc; }
TAILN=14 step tc_conv_bench 200 python tools/tc_conv_bench.py
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
export BNET_BENCH_STACKS=100 BNET_BENCH_LOG_DIR=$PWD/$OUT/arms
TAILN=14 step arm_stock 170 $TR --master-port 29631 bench.py --gpus $NG --steps 10 --warmup 3 --comm nccl --no-e2e --no-arms
TAILN=14 step arm_plugin 170 $TR --master-port 29632 bench.py --gpus $NG --steps 10 --warmup 3 --comm nccl-plugin --no-e2e --no-arms
TAILN=14 step arm_plugin_nograph 170 $TR --master-port 29633 bench.py --gpus $NG --steps 10 --warmup 3 --comm nccl-plugin --no-e2e --no-arms --no-graph
TAILN=16 step bench2 520 $TR --master-port 29634 bench.py --gpus $NG --steps 20 --warmup 5
TAILN=12 step gpu_tests_new 600 python -m pytest tests/test_gpu.py -q -x -k "collectives_on_ordinary or tcgen05 or staggered or transport_ring or allreduce_kernels or smoke"
echo "== done $(date -u)"
