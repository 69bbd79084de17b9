#!/bin/bash
# NCCL-over-the-plugin throughput sweep (2 GPUs is enough): executor geometry x NCCL pipeline settings.
# usage (on the GPU box): tools/gpu_plugin_perf.sh [tag] [nranks]
TAG=${1:-plugperf}; NG=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD
exec > >(tee $OUT/session.log) 2>&1
make -j16 >/dev/null && make bench >/dev/null || exit 1
BASE="$(python -m bagua_net_b200.utils.env)"
TUNED="$(python -m bagua_net_b200.utils.env --tuned)"
run() { local name=$1; shift; echo "---- [$name] $*"; timeout -k 5 120 env "$@" build/bench/all_reduce_perf -b 64K -e 256M -f 4 -N $NG -d bfloat16 -n 10 -w 3 > $OUT/$name.log 2>&1; echo "---- [$name] rc=$?"; grep -v "^$\|^#" $OUT/$name.log | tail -7 | cut -c1-120; }
run stock_nccl
run default            $BASE
run tuned              $TUNED
run tuned_chunk64k     $TUNED BNET_DEV_MIN_CHUNKSIZE=65536
run tuned_chunk1m      $TUNED BNET_DEV_MIN_CHUNKSIZE=1048576
run tuned_16clusters   $TUNED BNET_NCLUSTERS=16 BNET_CLUSTER_SIZE=2
run tuned_tma          $TUNED BNET_COPY_ENGINE=tma
run tuned_ce           $TUNED BNET_COPY_ENGINE=ce
run tuned_grid         $TUNED BNET_EXEC_GRID=1
run tuned_hostdirect   $TUNED BNET_HOST_SRC_DIRECT=1 NCCL_PROTO=LL,LL128,Simple
echo "== done"
