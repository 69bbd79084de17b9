// What a blocking poll loop does while nothing moves.  The loops of the transport collectives (one polling thread per rank
// drives isend / irecv / test) poll hot while requests complete: on a GPU box the next completion is microseconds away and
// the core is the thread's own.  Only a stretch without any progress backs off, in two stages chosen on an oversubscribed
// host (8 processes + the TCP transports' worker threads on 8 cores: 153 ms -> 39 ms for a 32 MiB all-reduce):
//   more than 64 idle passes           sched_yield() per pass — a no-op syscall when nobody else wants the core
//   more than 1 ms without progress    30 us sleeps: the peer is late by a scheduler's time scale, not by a kernel's
#pragma once
#include <sched.h>
#include <unistd.h>

#include "core/common.h"

namespace bnet {

struct IdleBackoff {
  unsigned idle = 0;
  uint64_t since = 0;
  void step(bool moved) {
    if (moved) {
      idle = 0;
      since = 0;
      return;
    }
    if (++idle <= 64) return;
    if ((idle & 63) == 0 || since) {
      const uint64_t now = now_ns();
      if (!since) since = now;
      else if (now - since > 1000000ull) {
        usleep(30);
        return;
      }
    }
    sched_yield();
  }
};

}  // namespace bnet
