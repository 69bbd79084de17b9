#!/bin/bash
# CPU-only sweep of the host transports through the ncclNet v8 table (bench/net_perf.cc):
# thread-per-stream TCP (BASIC), epoll TCP (TOKIO) for 1/2/4/8 streams, and the shared-memory ring.
cd "$(dirname "$0")/.."
make -s build/bench/net_perf || exit 1
B=build/bench/net_perf
ARGS="-b 1K -e 64M -f 4 -t ${BYTES:-2e9}"
echo "## $(nproc) cores, $(uname -r), loopback interface; window 8 requests; $(date -u +%F)"
for impl in BASIC TOKIO; do
  for ns in 1 2 4 8; do
    echo "### BNET_NVL=0 BAGUA_NET_IMPLEMENT=$impl BAGUA_NET_NSTREAMS=$ns"
    BNET_NVL=0 BAGUA_NET_IMPLEMENT=$impl BAGUA_NET_NSTREAMS=$ns timeout 300 $B $ARGS
  done
done
echo "### BNET_NVL=1 (same-host shared-memory ring, host buffers)"
BNET_NVL=1 timeout 300 $B $ARGS
