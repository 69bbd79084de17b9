#!/bin/bash
# Round-2 session 7 (8 GPUs, short — every minute costs eight): NCCL over the plugin at 8 ranks (nccl-tests sweep), the full
# bench with both DDP arms, stock NCCL sweep last.
TAG=${1:-r2s7}; NG=${2:-8}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD
exec > >(tee $OUT/session.log) 2>&1
echo "== session $TAG ngpus=$NG $(date -u)"
make -j16 >/dev/null 2>&1; make bench >/dev/null 2>&1
BASE="$(python -m bagua_net_b200.utils.env) BNET_WATCHDOG_MS=5000 NCCL_DEBUG=WARN"
ARP="build/bench/all_reduce_perf -N $NG -d bfloat16"
SWEEP="-b 8 -e 128M -f 4 -n 10 -w 3"
echo "---- [plugin] $(date -u +%T)"; timeout -k 5 100 env $BASE $ARP $SWEEP > $OUT/plugin.log 2>&1; echo "---- [plugin] rc=$?"; grep -v "^$" $OUT/plugin.log | tail -16 | cut -c1-200
export BNET_BENCH_STACKS=100 BNET_BENCH_LOG_DIR=$PWD/$OUT/arms
echo "---- [bench_full] $(date -u +%T)"
timeout -k 5 ${BENCH_TMO:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29634 bench.py --gpus $NG --steps 20 --warmup 5 --arm-timeout ${ARM_TMO:-90} > $OUT/bench_full.log 2>&1; echo "---- [bench_full] rc=$?"
grep -v "Warning\|warn\|^$\|\*\*\*\|OMP_NUM\|^  File\|^Thread\|no Python frame" $OUT/bench_full.log | tail -14 | cut -c1-3800
echo "---- [plugin_info] $(date -u +%T)"; timeout -k 5 40 env $BASE NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING $ARP -b 8 -e 64K -f 64 -n 5 -w 2 > $OUT/plugin_info.log 2>&1; echo "rc=$?"
grep -i "Using network\|via NET\|tuner" $OUT/plugin_info.log | sed 's/.*NCCL INFO //' | sed 's/[0-9]*\[[0-9]*\] -> [0-9]*\[[0-9]*\]/A->B/; s/Channel [0-9]*\/[0-9]*/Channel/; s/BNet\/[0-9]/BNet\/x/' | sort | uniq -c | sort -rn | head -6 | cut -c1-200
if [ -n "$WITH_STOCK" ]; then echo "---- [stock] $(date -u +%T)"; timeout -k 5 100 $ARP $SWEEP > $OUT/stock.log 2>&1; echo "---- [stock] rc=$?"; tail -14 $OUT/stock.log | cut -c1-200; fi
echo "== done $(date -u)"
