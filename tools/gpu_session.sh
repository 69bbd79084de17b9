#!/bin/bash
# One entry point for every measurement session on a gpurun box.  A session is a list of RECIPES run in order; each recipe
# is bounded by its own timeout, logs to gpurun_out/<tag>/<step>.log and prints the tail, so that a box cut short still
# leaves what finished.  What is worth keeping is copied to profiles/ by hand afterwards (tools/summarize_ncu.sh for
# .ncu-rep files).
#
#   gpurun [--gpus N] --timeout S -- 'bash tools/gpu_session.sh <tag> <ngpus> <recipe> [<recipe> ...]'
#
# recipes (the sessions behind profiles/README.md were these lists):
#   tests        pytest -m gpu + __graft_entry__.smoke()                         (1+ GPUs)
#   bench        bench.py with both DDP arms at N GPUs (N = 1: the plain run)    r2s5-r2s7: "plugin coll bench"
#   resnet       bench.py --model resnet50 --no-arms
#   plugin       nccl-tests clone over the plugin: 8 B .. 128 MiB, then a short INFO run (network / GDRDMA / tuner lines)
#   stock        the same sweep with stock NCCL (no plugin)
#   knobs        the 32 MiB / 128 MiB rows under env variants (KNOBS="A=1,B=2;C=3": one variant per ';')
#   collnet      the CollNet experiment: nccl-tests clone with BNET_COLLNET=1 NCCL_COLLNET_ENABLE=1, one virtual host per rank
#                (NCCL then sees one GPU per "node" and may pick CollNetDirect / CollNetChain: grep "CollNet" in collnet_info.log)
#   debug        bounded small sweeps over the plugin with the watchdog at 3 s and the transport's INFO log
#   coll         symmetric-heap collectives (LL latency, P2P / NVLS bandwidth), transport ring, staggered fused SGD
#   sweep        bench/allreduce_sweep.py: fused kernels vs stock NCCL, 1 KiB .. 1 GiB
#   tc           tcgen05 probes, linear vs cuBLAS, conv vs cuDNN, ncu of the linear kernel
#   kernels      layer-kernel table (tools/nn_kernel_bench.py) + ncu of the backward kernels
#   ncu-p2p      NVLink-side ncu of the per-message executor kernel, one process / two GPUs (needs N >= 2)
#   ncu-step     launch list of one training step + ncu of the fused all-reduce/SGD and layer kernels (1 GPU)
#   sanitize     compute-sanitizer memcheck + racecheck: executor, fused SGD (1 GPU); memcheck all-reduce (N >= 2)
#   sanitize-all adds synccheck, the fused layer kernels, pack/cast and the tcgen05 kernels (slow: ~15 min)
TAG=${1:?tag}; NG=${2:?ngpus}; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD
exec > >(tee -a $OUT/session.log) 2>&1
echo "== session $TAG ngpus=$NG recipes: $* ($(date -u))"
make -j16 >/dev/null 2>&1; make bench >/dev/null 2>&1

NOISE="Warning\|warn\|^$\|\*\*\*\|OMP_NUM\|return func\|return Variable\|^  File\|^Thread\|no Python frame"
step() {   # step <name> <timeout> <cmd...>   (TAILN / CUT tune what is echoed)
  local name=$1 tmo=$2; shift 2
  echo "---- [$name] $(date -u +%T)"
  timeout -k 5 $tmo "$@" > $OUT/$name.log 2>&1; local rc=$?
  echo "---- [$name] rc=$rc"; grep -v "$NOISE" $OUT/$name.log | tail -${TAILN:-8} | cut -c1-${CUT:-400}
  return $rc
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
PLUG="$(python -m bagua_net_b200.utils.env) BNET_WATCHDOG_MS=5000 NCCL_DEBUG=WARN"
ARP="build/bench/all_reduce_perf -N $NG -d bfloat16"
SWEEP="-b 8 -e 128M -f 4 -n 10 -w 3"
CS="compute-sanitizer --error-exitcode 66 --launch-timeout 0"
san() {    # san <name> <timeout> <cmd...>
  local name=$1 tmo=$2; shift 2
  echo "---- [$name] $(date -u +%T)"
  timeout -k 5 $tmo "$@" > $OUT/$name.log 2>&1; local rc=$?
  echo "---- [$name] rc=$rc  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/$name.log | tail -1)"
  grep -E "Invalid|Race reported|hazard|Barrier error|misaligned" $OUT/$name.log | head -4 | cut -c1-250
}

for recipe in "$@"; do
  case $recipe in
  tests)
    TAILN=15 step pytest_gpu 1500 python -m pytest tests -m gpu -x -q
    step smoke 300 python -c "import __graft_entry__ as g; g.smoke()" ;;
  bench)
    # (N > 1: the run also spawns the DDP arms, the transport-collective child, the CollNet probe and the ResNet arms: ~4 min at 8 GPUs)
    export BNET_BENCH_STACKS=${BNET_BENCH_STACKS:-140} BNET_BENCH_LOG_DIR=$PWD/$OUT/arms
    if [ "$NG" = 1 ]; then TAILN=3 CUT=4000 step bench 400 python bench.py --steps 20 --warmup 5
    else TAILN=14 CUT=6000 step bench 700 $TR --master-port 29634 bench.py --gpus $NG --steps 20 --warmup 5 --arm-timeout ${ARM_TMO:-150}; fi
    unset BNET_BENCH_STACKS ;;
  resnet)
    if [ "$NG" = 1 ]; then TAILN=3 CUT=4000 step bench_resnet50 400 python bench.py --steps 20 --warmup 5 --model resnet50
    else TAILN=4 CUT=4000 step bench_resnet50 400 $TR --master-port 29635 bench.py --gpus $NG --steps 20 --warmup 5 --model resnet50 --no-arms; fi ;;
  plugin)
    # (every process of a fresh box needs ~20 s before its first line at 8 ranks: keep these timeouts generous)
    TAILN=16 step plugin 150 env $PLUG $ARP $SWEEP
    TAILN=2 step plugin_info 90 env $PLUG NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING BNET_LOG_LEVEL=2 $ARP -b 8 -e 64K -f 64 -n 5 -w 2
    grep -i "Using network\|via NET\|tuner" $OUT/plugin_info.log | sed 's/.*NCCL INFO //; s/[0-9]*\[[0-9]*\] -> [0-9]*\[[0-9]*\]/A->B/; s/Channel [0-9]*\/[0-9]*/Channel/; s/BNet\/[0-9]/BNet\/x/' | sort | uniq -c | sort -rn | head -6 | cut -c1-200 ;;
  stock)
    TAILN=16 step stock 150 $ARP $SWEEP ;;
  collnet)
    CN="BNET_COLLNET=1 NCCL_COLLNET_ENABLE=1 NCCL_COLLNET_NODE_THRESHOLD=1"
    TAILN=6 step collnet_info 120 env $PLUG $CN NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,NET,COLL,TUNING BNET_LOG_LEVEL=2 $ARP -H -b 1M -e 16M -f 4 -n 5 -w 2
    grep -i "collnet\|coll net" $OUT/collnet_info.log | sed 's/.*NCCL INFO //' | sort | uniq -c | sort -rn | head -8 | cut -c1-200
    TAILN=16 step collnet 150 env $PLUG $CN $ARP -H $SWEEP
    TAILN=16 step collnet_forced 150 env $PLUG $CN NCCL_ALGO=CollnetDirect,CollnetChain,Ring $ARP -H $SWEEP ;;
  knobs)
    i=0; IFS=';' read -ra VARIANTS <<< "${KNOBS:-BNET_MSG_BATCH_US=20;NCCL_MIN_NCHANNELS=8,NCCL_MAX_NCHANNELS=8;NCCL_BUFFSIZE=8388608;BNET_MSG_CLUSTER=0;BNET_EXEC_MODE=ce}"
    TAILN=4 step knob_base 90 env $PLUG BNET_EXEC_STATS=1 $ARP -b 32M -e 128M -f 4 -n 8 -w 2
    for v in "${VARIANTS[@]}"; do i=$((i+1)); echo "variant $i: $v"; TAILN=4 step knob_$i 90 env $PLUG BNET_EXEC_STATS=1 ${v//,/ } $ARP -b 32M -e 128M -f 4 -n 8 -w 2; done ;;
  debug)
    DBG="$PLUG BNET_WATCHDOG_MS=3000 BNET_LOG_LEVEL=2"
    TAILN=12 step debug_small 120 env $DBG $ARP -b 8 -e 1M -f 8 -n 3 -w 1
    TAILN=12 step debug_hostptr 120 env $DBG BNET_GDR=0 $ARP -b 8 -e 1M -f 8 -n 3 -w 1
    TAILN=12 step debug_persistent 120 env $DBG BNET_EXEC_MODE=persistent $ARP -b 8 -e 1M -f 8 -n 3 -w 1 ;;
  coll)
    TAILN=12 step coll_allreduce 240 $TR --master-port 29641 tests/gpu_worker.py allreduce
    TAILN=6 step transport_ring 200 $TR --master-port 29642 tests/gpu_worker.py transport_ring
    TAILN=4 step fused_sgd_staggered 150 $TR --master-port 29643 tests/gpu_worker.py fused_sgd_staggered
    TAILN=4 step transport_mesh 150 $TR --master-port 29645 tests/gpu_worker.py transport_mesh
    TAILN=4 step transport_mesh_twoshot 150 $TR --master-port 29647 tests/gpu_worker.py transport_mesh_twoshot
    TAILN=4 step transport_ring_compressed 150 $TR --master-port 29646 tests/gpu_worker.py transport_ring_compressed ;;
  sweep)
    TAILN=30 step allreduce_sweep 400 $TR --master-port 29644 bench/allreduce_sweep.py ;;
  tc)
    TAILN=30 step tc_probe 300 python tools/tc_probe.py
    TAILN=20 step tc_linear_bench 300 python tools/tc_linear_bench.py
    TAILN=24 step tc_conv_bench 300 python tools/tc_conv_bench.py --wgrad
    step ncu_tc 400 ncu --set full --clock-control none --import-source on -k "regex:tc_linear_kernel" -c 4 -f -o $OUT/tc_linear \
      python tools/tc_linear_bench.py --iters 1 --warmup 0 --shapes 4096x4096x4096,32x4096x25088
    # the convolution kernels at the conv4 shape: launch 0 = forward, 1 = input gradient, 2.. = filter gradient (BNET_TC_WGRAD=1: no child check)
    step ncu_tc_conv 300 env BNET_TC_WGRAD=1 ncu --set full --clock-control none --import-source on -k regex:tc_linear_kernel -c 4 -f -o $OUT/tc_conv \
      python tools/tc_conv_bench.py --wgrad --iters 1 --shapes 512x512x28 ;;
  kernels)
    TAILN=12 step nn_kernel_bench 200 python tools/nn_kernel_bench.py
    step ncu_nn 400 ncu --set full --clock-control none --import-source on -k "regex:relu_bwd_bias_grad|pool_relu_bwd" -c 6 -f -o $OUT/nn_bwd \
      python tools/nn_kernel_bench.py --iters 1 --warmup 0 ;;
  ncu-p2p)
    TAILN=8 step p2p_time 150 python tools/ncu_p2p.py --time
    # launch 0 of the filtered kernel is the executor's own warm-up; 1..3 = copy, red.add f32, bf16 -> f32 accumulate
    step ncu_p2p 240 ncu --set full --clock-control none --import-source on -k regex:bnet_nvl_msg --launch-skip 1 -c 3 -f -o $OUT/nvl_msg \
      python tools/ncu_p2p.py --mbytes 128
    ncu -i $OUT/nvl_msg.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_l1tex2xbar_write_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,nvltx__bytes.sum,nvlrx__bytes.sum,launch__grid_size,launch__registers_per_thread,sm__warps_active.avg.per_cycle_active \
      > $OUT/nvl_msg.raw.csv 2>/dev/null; cut -c1-600 $OUT/nvl_msg.raw.csv ;;
  ncu-step)
    step ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file $OUT/launches.csv \
      python bench.py --gpus 1 --steps 2 --warmup 3 --no-e2e
    step ncu_fused 400 ncu --set full --clock-control none --import-source on -k regex:bnet_fused -s 6 -c 2 -f -o $OUT/fused_sgd \
      python bench.py --gpus 1 --steps 2 --warmup 3 --no-e2e
    step ncu_nn_step 400 ncu --set full --clock-control none --import-source on -k regex:bnet::nn -s 40 -c 8 -f -o $OUT/nn_kernels \
      python bench.py --gpus 1 --steps 2 --warmup 3 --no-e2e ;;
  sanitize|sanitize-all)
    export BNET_TEST_QUICK=1
    tools="memcheck racecheck"; [ $recipe = sanitize-all ] && tools="memcheck racecheck synccheck"
    for t in $tools; do
      san ${t}_executor_msg 200 env BNET_EXEC_MODE=msg $CS --tool $t python tests/gpu_worker.py executor
      san ${t}_fused_sgd 200 $CS --tool $t python tests/gpu_worker.py fused_sgd
      if [ $recipe = sanitize-all ]; then
        san ${t}_executor_persistent 300 env BNET_EXEC_MODE=persistent $CS --tool $t python tests/gpu_worker.py executor
        san ${t}_fused_nn 400 $CS --tool $t python tests/gpu_worker.py fused_nn
        san ${t}_pack_cast 200 $CS --tool $t python tests/gpu_worker.py pack_cast
      fi
    done
    if [ $recipe = sanitize-all ]; then
      san memcheck_tc_linear 400 $CS --tool memcheck python tests/gpu_worker.py tc_linear
      san memcheck_tc_conv 400 $CS --tool memcheck python tests/gpu_worker.py tc_conv
      san synccheck_tc_conv 400 $CS --tool synccheck python tests/gpu_worker.py tc_conv
    fi
    if [ "$NG" -ge 2 ]; then
      TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
      san memcheck_allreduce_2gpu 300 $CS --tool memcheck --target-processes all $TR2 --master-port 29655 tests/gpu_worker.py allreduce
      [ $recipe = sanitize-all ] && san racecheck_fused_sgd_2gpu 400 $CS --tool racecheck --target-processes all $TR2 --master-port 29657 tests/gpu_worker.py fused_sgd
    fi
    unset BNET_TEST_QUICK
    echo "== sanitizer summary"; grep -H -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/*check*.log | sed "s#$OUT/##" ;;
  *) echo "unknown recipe '$recipe'"; exit 2 ;;
  esac
done
echo "== done $(date -u)"
