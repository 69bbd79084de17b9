#!/bin/bash
# Round-2 session 4 (N GPUs): tuner + batching window + EAGER module loading; the DDP arms; the full bench with arms.
TAG=${1:-r2s4}; NG=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD
exec > >(tee $OUT/session.log) 2>&1
echo "== session $TAG ngpus=$NG $(date -u)"
make -j16 >/dev/null 2>&1; make bench >/dev/null 2>&1
BASE="$(python -m bagua_net_b200.utils.env) BNET_WATCHDOG_MS=5000 NCCL_DEBUG=WARN"
echo "env: $BASE" | cut -c1-600
ARP="build/bench/all_reduce_perf -N $NG -d bfloat16"
run() { local name=$1 tmo=$2 pargs=$3; shift 3; echo "---- [$name] $(date -u +%T) $*"; timeout -k 5 $tmo env $BASE "$@" $ARP $pargs > $OUT/$name.log 2>&1; local rc=$?; echo "---- [$name] rc=$rc"; grep -v "^$" $OUT/$name.log | tail -${TAILN:-16} | cut -c1-330; return $rc; }
SWEEP="-b 8 -e 128M -f 4 -n 10 -w 3"
echo "---- [stock]"; timeout -k 5 60 $ARP $SWEEP > $OUT/stock.log 2>&1; echo "rc=$?"; tail -14 $OUT/stock.log
run tuner_batched 60 "$SWEEP"
TAILN=6 run tuner_info 40 "-b 8 -e 64K -f 64 -n 5 -w 2" NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING
grep -i "tuner" $OUT/tuner_info.log | head -5 | cut -c1-200
run no_tuner 60 "$SWEEP" NCCL_TUNER_PLUGIN=none BNET_TUNER=0
step() { local name=$1 tmo=$2; shift 2; echo "---- [$name] $(date -u +%T)"; timeout -k 5 $tmo "$@" > $OUT/$name.log 2>&1; local rc=$?; echo "---- [$name] rc=$rc"; grep -v "Warning\|warn\|^$\|\*\*\*\|OMP_NUM\|return func\|return Variable" $OUT/$name.log | tail -${TAILN:-8} | cut -c1-${CUT:-3000}; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
export BNET_BENCH_STACKS=100 BNET_BENCH_LOG_DIR=$PWD/$OUT/arms
TAILN=12 step arm_plugin 170 $TR --master-port 29632 bench.py --gpus $NG --steps 10 --warmup 3 --comm nccl-plugin --no-e2e --no-arms
TAILN=12 step arm_stock 120 $TR --master-port 29631 bench.py --gpus $NG --steps 10 --warmup 3 --comm nccl --no-e2e --no-arms
TAILN=14 step bench_full 420 $TR --master-port 29634 bench.py --gpus $NG --steps 20 --warmup 5
if [ "$3" = "resnet" ]; then TAILN=6 step bench_resnet50 420 $TR --master-port 29635 bench.py --gpus $NG --steps 20 --warmup 5 --model resnet50; fi
echo "== done $(date -u)"
