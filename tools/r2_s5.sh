#!/bin/bash
# Round-2 session 5 (N = 4 or 8 GPUs): NCCL over the plugin at N ranks (nccl-tests sweep vs stock), the collectives'
# latency / bandwidth lines, the all-reduce that rides the transport, the full bench with both DDP arms.
TAG=${1:-r2s5}; NG=${2:-4}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD
exec > >(tee $OUT/session.log) 2>&1
echo "== session $TAG ngpus=$NG $(date -u)"
make -j16 >/dev/null 2>&1; make bench >/dev/null 2>&1
ls -la bagua_net_b200/lib | grep tuner
BASE="$(python -m bagua_net_b200.utils.env) BNET_WATCHDOG_MS=5000 NCCL_DEBUG=WARN"
ARP="build/bench/all_reduce_perf -N $NG -d bfloat16"
run() { local name=$1 tmo=$2 pargs=$3; shift 3; echo "---- [$name] $(date -u +%T) $*"; timeout -k 5 $tmo env $BASE "$@" $ARP $pargs > $OUT/$name.log 2>&1; local rc=$?; echo "---- [$name] rc=$rc"; grep -v "^$" $OUT/$name.log | tail -${TAILN:-16} | cut -c1-330; return $rc; }
SWEEP="-b 8 -e 128M -f 4 -n 10 -w 3"
echo "---- [stock]"; timeout -k 5 60 $ARP $SWEEP > $OUT/stock.log 2>&1; echo "rc=$?"; tail -14 $OUT/stock.log
run plugin 90 "$SWEEP"
TAILN=4 run plugin_info 40 "-b 8 -e 64K -f 64 -n 5 -w 2" NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING BNET_LOG_LEVEL=2
grep -i "tuner\|Using network\|via NET" $OUT/plugin_info.log | sort | uniq -c | sort -rn | head -8 | cut -c1-220
step() { local name=$1 tmo=$2; shift 2; echo "---- [$name] $(date -u +%T)"; timeout -k 5 $tmo "$@" > $OUT/$name.log 2>&1; local rc=$?; echo "---- [$name] rc=$rc"; grep -v "Warning\|warn\|^$\|\*\*\*\|OMP_NUM\|return func\|return Variable" $OUT/$name.log | tail -${TAILN:-8} | cut -c1-${CUT:-3500}; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
TAILN=12 step coll_allreduce 200 $TR --master-port 29641 tests/gpu_worker.py allreduce
TAILN=6 step transport_ring 150 $TR --master-port 29642 tests/gpu_worker.py transport_ring
TAILN=4 step fused_sgd_staggered 120 $TR --master-port 29643 tests/gpu_worker.py fused_sgd_staggered
export BNET_BENCH_STACKS=280 BNET_BENCH_LOG_DIR=$PWD/$OUT/arms
TAILN=14 step bench_full 640 $TR --master-port 29634 bench.py --gpus $NG --steps 20 --warmup 5
if [ "$3" = "resnet" ]; then TAILN=4 step bench_resnet50 300 $TR --master-port 29635 bench.py --gpus $NG --steps 20 --warmup 5 --model resnet50 --no-arms; fi
echo "== done $(date -u)"
