// net_perf — point-to-point bandwidth / message-rate of the transport behind the ncclNet table,
// measured the way NCCL's proxy drives it: a window of up to 8 requests in flight per connection
// (NCCL_NET_MAX_REQUESTS, reference: cc/nccl_types.h:50), isend/irecv/test polled from one thread
// per side, sender and receiver in two processes.  No GPU: host buffers over loopback TCP or the
// shared-memory ring — this is the reference-equivalent data path (reference:
// src/implement/nthread_per_socket_backend.rs:524-631).
//
//   net_perf [-l libnccl-net.so] [-b min] [-e max] [-f factor] [-w window] [-t bytes moved per size, default 2e9] [-c check]
//            [-m host|fakecuda]
// honours the BAGUA_NET_* / BNET_* environment (implementation, streams, chunk size, NVL on/off).
// Output columns: bytes, messages, time, GB/s, messages/s, mean us per message, longest gap between completions (@ message).
#include <dlfcn.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <vector>

#include "bnet/nccl_net_abi.h"

static double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static size_t parse_size(const char* s) {
  char* end;
  double v = strtod(s, &end);
  switch (*end) {
    case 'k': case 'K': v *= 1024; break;
    case 'm': case 'M': v *= 1024.0 * 1024; break;
    case 'g': case 'G': v *= 1024.0 * 1024 * 1024; break;
  }
  return (size_t)v;
}

#define OK(call)                                                          \
  do {                                                                    \
    ncclResult_t r_ = (call);                                             \
    if (r_ != ncclSuccess) { fprintf(stderr, "%s:%d %s -> %d\n", __FILE__, __LINE__, #call, (int)r_); _exit(3); } \
  } while (0)

struct Shared {   // cross-process control block
  volatile int errors;
};

int main(int argc, char** argv) {
  const char* lib = "bagua_net_b200/lib/libnccl-net.so";
  size_t lo = 64, hi = 64u << 20;
  double factor = 4, budget = 2e9;
  int window = 8, check = 1;
  bool fake = false;
  for (int i = 1; i + 1 < argc; i += 2) {
    if (!strcmp(argv[i], "-l")) lib = argv[i + 1];
    else if (!strcmp(argv[i], "-b")) lo = parse_size(argv[i + 1]);
    else if (!strcmp(argv[i], "-e")) hi = parse_size(argv[i + 1]);
    else if (!strcmp(argv[i], "-f")) factor = atof(argv[i + 1]);
    else if (!strcmp(argv[i], "-w")) window = atoi(argv[i + 1]);
    else if (!strcmp(argv[i], "-t")) budget = (double)parse_size(argv[i + 1]);
    else if (!strcmp(argv[i], "-c")) check = atoi(argv[i + 1]);
    else if (!strcmp(argv[i], "-m")) fake = !strcmp(argv[i + 1], "fakecuda");
    else { fprintf(stderr, "unknown option %s\n", argv[i]); return 2; }
  }
  if (window < 1) window = 1;
  if (window > 8) window = 8;
  if (factor < 1.01) factor = 2;

  Shared* sh = (Shared*)mmap(nullptr, 4096, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  memset((void*)sh, 0, sizeof(*sh));
  int hpipe[2];
  if (pipe(hpipe)) return 2;

  pid_t pid = fork();
  const bool sender = pid == 0;
  void* h = dlopen(lib, RTLD_NOW);
  if (!h) { fprintf(stderr, "dlopen(%s): %s\n", lib, dlerror()); return 2; }
  ncclNet_v8_t* net = (ncclNet_v8_t*)dlsym(h, "ncclNetPlugin_v8");
  if (!net) { fprintf(stderr, "ncclNetPlugin_v8 missing\n"); return 2; }
  OK(net->init(nullptr));
  char handle[NCCL_NET_HANDLE_MAXSIZE] = {0};
  void *lcomm = nullptr, *comm = nullptr;
  if (!sender) {
    OK(net->listen(0, handle, &lcomm));
    if (write(hpipe[1], handle, sizeof(handle)) != (ssize_t)sizeof(handle)) return 2;
    while (!comm) OK(net->accept(lcomm, &comm, nullptr));
  } else {
    if (read(hpipe[0], handle, sizeof(handle)) != (ssize_t)sizeof(handle)) _exit(2);
    while (!comm) OK(net->connect(0, handle, &comm, nullptr));
  }

  // -m fakecuda (with BNET_FAKE_CUDA=1): "device" buffers emulated by shm segments — the registered-buffer /
  // descriptor / completion-word protocol of the NVLink direct path with a memcpy standing in for the kernel,
  // i.e. the host-side cost per message of that path
  struct Buf {
    unsigned char* p = nullptr;
    std::vector<unsigned char> own;
    unsigned char* data() { return p; }
    unsigned char& operator[](size_t i) { return p[i]; }
  } buf[8];
  void* mh[8];
  typedef void* (*fake_alloc_t)(size_t);
  fake_alloc_t fake_alloc = fake ? (fake_alloc_t)dlsym(h, "bnet_fake_cuda_alloc") : nullptr;
  if (fake && !fake_alloc) { fprintf(stderr, "bnet_fake_cuda_alloc missing\n"); return 2; }
  for (int j = 0; j < window; j++) {
    const size_t cap = hi + 64;
    if (fake) {
      buf[j].p = (unsigned char*)fake_alloc(cap);
      if (!buf[j].p) { fprintf(stderr, "emulated device allocation failed (BNET_FAKE_CUDA=1 set?)\n"); return 2; }
    } else {
      buf[j].own.resize(cap);
      buf[j].p = buf[j].own.data();
    }
    for (size_t k = 0; k < cap; k += 61) buf[j][k] = sender ? (unsigned char)(k * 7 + 3) : 0;
    OK(net->regMr(comm, buf[j].data(), cap, fake ? NCCL_PTR_CUDA : NCCL_PTR_HOST, &mh[j]));
  }
  if (!sender) printf("# %-10s %10s %9s %9s %12s %10s %11s\n", "bytes", "messages", "time(s)", "GB/s", "messages/s", "us/msg", "max gap us");

  for (double fs = (double)lo; (size_t)fs <= hi; fs *= factor) {
    const size_t size = (size_t)fs;
    // both sides derive the same message count: ~`budget` bytes per size, bounded for tiny / huge messages
    long long count = (long long)(budget / (double)(size ? size : 1));
    if (count < 32) count = 32;
    if (count > 200000) count = 200000;
    // one untimed round trip so both sides start the size together
    for (int warm = 0; warm < 2; warm++) {
      void* r = nullptr;
      if (sender) {
        while (!r) OK(net->isend(comm, buf[0].data(), 8, 1, mh[0], &r));
      } else {
        void* d = buf[0].data();
        int cap = 64, tag = 1;
        while (!r) OK(net->irecv(comm, 1, &d, &cap, &tag, &mh[0], &r));
      }
      int done = 0;
      while (!done) OK(net->test(r, &done, nullptr));
    }
    void* req[8] = {nullptr};
    long long posted = 0, completed = 0, gap_at = 0;
    const double t0 = now_s();
    double last_done = t0, max_gap = 0;   // longest wait between two completions: stalls hide in an average
    while (completed < count) {
      const long long before = posted + completed;
      for (int j = 0; j < window && posted < count; j++) {   // post while the window has room
        if (req[j]) continue;
        if (sender) {
          OK(net->isend(comm, buf[j].data(), (int)size, 1, mh[j], &req[j]));
        } else {
          void* d = buf[j].data();
          int cap = (int)(size + 64), tag = 1;
          OK(net->irecv(comm, 1, &d, &cap, &tag, &mh[j], &req[j]));
        }
        if (req[j]) posted++;
      }
      for (int j = 0; j < window; j++) {                      // reap
        if (!req[j]) continue;
        int done = 0, got = -1;
        OK(net->test(req[j], &done, &got));
        if (!done) continue;
        if ((size_t)got != size) sh->errors++;
        if (check && !sender && size > 61 && buf[j][61] != (unsigned char)(61 * 7 + 3)) sh->errors++;
        req[j] = nullptr;
        completed++;
        if ((completed & 15) == 0 || completed < 64) {
          const double t = now_s();
          if (t - last_done > max_gap) { max_gap = t - last_done; gap_at = completed; }
          last_done = t;
        }
      }
      // a pass that neither posted nor reaped: give the core away, as NCCL's proxy thread does when it is idle — two
      // spinning pollers that land on one core otherwise trade whole time slices (20 ms gaps, 1 ms per message)
      if (posted + completed == before) sched_yield();
    }
    const double dt = now_s() - t0;
    if (!sender) {
      printf("%-12zu %10lld %9.3f %9.3f %12.0f %10.2f %11.1f @%lld\n", size, completed, dt, completed * (double)size / dt / 1e9,
             completed / dt, dt / completed * 1e6, max_gap * 1e6, gap_at);
      fflush(stdout);
    }
  }
  for (int j = 0; j < window; j++) net->deregMr(comm, mh[j]);
  if (sender) { net->closeSend(comm); _exit(0); }
  net->closeRecv(comm);
  net->closeListen(lcomm);
  int st = 0;
  waitpid(pid, &st, 0);
  if (sh->errors) printf("# %d ERRORS\n", sh->errors);
  return (sh->errors || st) ? 1 : 0;
}
