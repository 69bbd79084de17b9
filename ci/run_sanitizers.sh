#!/bin/bash
# Host engine under ThreadSanitizer and Address/UB sanitizers (the reference has no sanitizer job).
# Builds build/{tsan,asan}/libnccl-net.so (`make tsan`, `make asan` run the C++ tests against them) and then
# drives the Python two-process loopback / ABI / telemetry tests against the same libraries through
# BNET_LIB_DIR + LD_PRELOAD of the sanitizer runtime.  Any sanitizer report fails the script.  The sanitizer libraries
# export every table (v3 .. v10, CollNet included), so they also stand in for the -bnet / -bnetx variants.
# (ODR detection is off: libnccl-net.so and its -bnetx variant export the same tables on purpose, and the doctor
# test loads both.)
set -e
cd "$(dirname "$0")/.."
GCCLIB=$(dirname "$(/usr/bin/g++ -print-file-name=libtsan.so)")
make -j"$(nproc)"
make emu-asan            # kernel bodies on an emulated grid under ASan/UBSan
for san in tsan asan; do
  make $san
  for alias in bnet bnetx; do cp -f build/$san/libnccl-net.so build/$san/libnccl-net-$alias.so; done
  logs=$(mktemp -d)
  if [ $san = tsan ]; then
    pre="$GCCLIB/libtsan.so"; export TSAN_OPTIONS="log_path=$logs/r exitcode=0 report_signal_unsafe=0"
  else
    pre="$GCCLIB/libasan.so $GCCLIB/libubsan.so"
    export ASAN_OPTIONS="detect_leaks=0 detect_odr_violation=0 log_path=$logs/r halt_on_error=0" UBSAN_OPTIONS="log_path=$logs/r print_stacktrace=1"
  fi
  BNET_LIB_DIR=$PWD/build/$san LD_PRELOAD="$pre" \
    python -m pytest tests/test_loopback.py tests/test_utils.py tests/test_telemetry.py -q -x
  if ls $logs/r* >/dev/null 2>&1; then echo "== $san reports:"; head -80 $logs/r*; exit 1; fi
  echo "== $san: clean"
done
