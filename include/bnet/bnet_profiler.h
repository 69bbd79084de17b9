/* bnet events in NCCL's profiler (ncclNet v10 `init(logFn, profFn)` / `isend(..., phandle, ...)`).
 *
 * NCCL >= 2.26 hands a net plugin a callback and, per isend/irecv, the handle of the proxy-step event the
 * transfer belongs to.  bnet reports one event per request through it:
 *
 *     profFn(&eHandle, 0 (start), pHandle, BNET_PROFILER_PLUGIN_ID, &descr)      at isend / irecv
 *     profFn(&eHandle, 1 (stop),  NULL,    BNET_PROFILER_PLUGIN_ID, &descr)      when test() reports completion
 *
 * The plugin id carries a type in the upper and a version in the lower 16 bits, like NCCL's own IB / socket
 * plugins do (their types are 1 and 2).  A profiler plugin that wants to show bnet's events decodes `descr` with
 * this header; profilers that do not know the type ignore the events.  The reference has no equivalent: its spans
 * go to Jaeger only (reference nthread_per_socket_backend.rs:529-538, 565-573).
 */
#ifndef BNET_PROFILER_H_
#define BNET_PROFILER_H_
#include <stddef.h>
#include <stdint.h>

#define BNET_PROFILER_NET_TYPE 0x42u                 /* 'B' */
#define BNET_PROFILER_NET_VERSION 1u
#define BNET_PROFILER_PLUGIN_ID (((int64_t)BNET_PROFILER_NET_TYPE << 16) | BNET_PROFILER_NET_VERSION)

enum { BNET_PROF_ISEND = 0, BNET_PROF_IRECV = 1 };
enum { BNET_PROF_PATH_TCP = 0, BNET_PROF_PATH_SHM_RING = 1, BNET_PROF_PATH_CMA = 2, BNET_PROF_PATH_NVLINK_KERNEL = 3 };

typedef struct {
  uint8_t type;          /* BNET_PROF_ISEND / BNET_PROF_IRECV */
  uint8_t path;          /* BNET_PROF_PATH_* (known at stop time; 0 at start) */
  uint16_t reserved;
  int32_t tag;
  uint64_t comm_id;      /* the connection (same id as in the Jaeger/OTLP span names isend-<comm>) */
  uint64_t request_id;
  size_t length;         /* bytes posted (start) / bytes moved (stop) */
} bnetProfilerEventDescr_v1_t;

#endif
