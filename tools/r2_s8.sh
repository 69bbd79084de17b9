#!/bin/bash
# Round-2 session 8 (2 GPUs, ~100 s): NVLink-side ncu of the per-message executor kernel (GPU0 -> GPU1, one process), and the
# layer-kernel / convolution tables after the atomics fix.
TAG=${1:-r2s8}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONPATH=$PWD
exec > >(tee $OUT/session.log) 2>&1
echo "== session $TAG $(date -u)"
make -j16 >/dev/null 2>&1
# launch 0 of the filtered kernel is the executor's warm-up; 1..3 = copy, red.add f32, bf16 -> f32 accumulate (128 MiB sources)
echo "---- [ncu_p2p] $(date -u +%T)"
timeout -k 5 150 ncu --set full --clock-control none --import-source on -k regex:bnet_nvl_msg --launch-skip 1 -c 3 -f -o $OUT/nvl_msg python tools/ncu_p2p.py --mbytes 128 > $OUT/ncu_p2p.log 2>&1; echo "rc=$?"; tail -4 $OUT/ncu_p2p.log | cut -c1-200
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes.sum.per_second,nvlrx__bytes.sum.per_second,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread,smsp__warps_active.avg.per_cycle_active"
ncu -i $OUT/nvl_msg.ncu-rep --page raw --csv --metrics $M > $OUT/nvl_msg.raw.csv 2>$OUT/nvl_msg.raw.err; cat $OUT/nvl_msg.raw.csv | cut -c1-900
echo "---- [nn_kernel_bench] $(date -u +%T)"
timeout -k 5 120 python tools/nn_kernel_bench.py > $OUT/nn_kernel_bench.log 2>&1; echo "rc=$?"; grep -v "^$\|Warn" $OUT/nn_kernel_bench.log | tail -40 | cut -c1-200
if [ -n "$WITH_CONV" ]; then
echo "---- [tc_conv_bench] $(date -u +%T)"
timeout -k 5 120 python tools/tc_conv_bench.py > $OUT/tc_conv_bench.log 2>&1; echo "rc=$?"; grep -v "^$\|Warn" $OUT/tc_conv_bench.log | tail -24 | cut -c1-200
fi
echo "== done $(date -u)"
