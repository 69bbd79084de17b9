// C++ loopback test (no GPU): dlopen the plugin, fetch ncclNetPlugin_v4 — the symbol the
// reference exports (reference: cc/v4/nccl_net_v4.cc:210) — and ncclNetPlugin_v8, and drive
// init -> devices -> getProperties -> listen -> connect -> accept -> regMr -> isend/irecv ->
// test -> close between two threads over loopback (BASELINE.json config #1).
// usage: loopback_test <path to libnccl-net.so>      (honours BAGUA_NET_* / BNET_* env)
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "bnet/nccl_net_abi.h"

static std::atomic<int> g_fail{0};
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) { printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); g_fail++; } \
  } while (0)

static void fill(std::vector<unsigned char>& b, unsigned seed) {
  unsigned x = seed * 2654435761u + 12345u;
  for (auto& c : b) { x = x * 1664525u + 1013904223u; c = (unsigned char)(x >> 24); }
}

template <typename NET, typename SendFn, typename RecvFn>
static void run_pair(NET* net, void* scomm, void* rcomm, SendFn isend, RecvFn irecv) {
  const size_t sizes[] = {0, 1, 8, 4096, 512 * 1024, (1 << 20) - 1, (1 << 20) + 1, 5 << 20};
  const int inflight = 8;
  for (size_t size : sizes) {
    std::vector<std::vector<unsigned char>> src(inflight), dst(inflight);
    std::vector<void*> sreq(inflight), rreq(inflight), smh(inflight), rmh(inflight);
    for (int j = 0; j < inflight; j++) {
      src[j].resize(size + 1);
      dst[j].assign(size + 64, 0);
      fill(src[j], (unsigned)(size * 31 + j));
      CHECK(net->regMr(scomm, src[j].data(), size, NCCL_PTR_HOST, &smh[j]) == ncclSuccess);
      CHECK(net->regMr(rcomm, dst[j].data(), size + 64, NCCL_PTR_HOST, &rmh[j]) == ncclSuccess);
    }
    std::thread rx([&] {
      for (int j = 0; j < inflight; j++) {
        rreq[j] = nullptr;
        while (!rreq[j]) CHECK(irecv(rcomm, dst[j].data(), (int)(size + 64), rmh[j], &rreq[j]) == ncclSuccess);
      }
      for (int j = 0; j < inflight; j++) {
        int done = 0, got = -1;
        while (!done) CHECK(net->test(rreq[j], &done, &got) == ncclSuccess);
        CHECK((size_t)got == size);
      }
    });
    for (int j = 0; j < inflight; j++) {
      sreq[j] = nullptr;
      while (!sreq[j]) CHECK(isend(scomm, src[j].data(), (int)size, smh[j], &sreq[j]) == ncclSuccess);
    }
    for (int j = 0; j < inflight; j++) {
      int done = 0, got = -1;
      while (!done) CHECK(net->test(sreq[j], &done, &got) == ncclSuccess);
      CHECK((size_t)got == size);
    }
    rx.join();
    for (int j = 0; j < inflight; j++) {
      CHECK(size == 0 || memcmp(src[j].data(), dst[j].data(), size) == 0);
      for (size_t k = size; k < size + 64; k++) CHECK(dst[j][k] == 0);
      net->deregMr(scomm, smh[j]);
      net->deregMr(rcomm, rmh[j]);
    }
  }
}

int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "bagua_net_b200/lib/libnccl-net.so";
  void* h = dlopen(path, RTLD_NOW);
  if (!h) { printf("dlopen(%s): %s\n", path, dlerror()); return 2; }
  {  // ---- v4: the reference's ABI (blocking accept)
    ncclNet_v4_t* net = (ncclNet_v4_t*)dlsym(h, "ncclNetPlugin_v4");
    if (!net) { printf("ncclNetPlugin_v4 missing\n"); return 2; }
    CHECK(!strcmp(net->name, "BNet"));
    CHECK(net->init(nullptr) == ncclSuccess);
    int ndev = 0;
    CHECK(net->devices(&ndev) == ncclSuccess && ndev >= 1);
    ncclNetProperties_v4_t props;
    CHECK(net->getProperties(0, &props) == ncclSuccess && props.name && (props.ptrSupport & NCCL_PTR_HOST) && props.maxComms > 0);
    char handle[NCCL_NET_HANDLE_MAXSIZE_V4];
    void *lcomm = nullptr, *scomm = nullptr, *rcomm = nullptr;
    CHECK(net->listen(0, handle, &lcomm) == ncclSuccess);
    std::thread acc([&] { CHECK(net->accept(lcomm, &rcomm) == ncclSuccess); });
    CHECK(net->connect(0, handle, &scomm) == ncclSuccess && scomm);
    acc.join();
    CHECK(rcomm);
    run_pair(net, scomm, rcomm,
             [&](void* c, void* d, int s, void* mh, void** r) { return net->isend(c, d, s, mh, r); },
             [&](void* c, void* d, int s, void* mh, void** r) { return net->irecv(c, d, s, mh, r); });
    void* fr = nullptr;
    char x;
    CHECK(net->iflush(rcomm, &x, 1, nullptr, &fr) == ncclSuccess);   // the reference errors here (nccl_net_v4.cc:145-149)
    if (fr) { int done = 0; while (!done) CHECK(net->test(fr, &done, nullptr) == ncclSuccess); }
    CHECK(net->closeSend(scomm) == ncclSuccess && net->closeRecv(rcomm) == ncclSuccess && net->closeListen(lcomm) == ncclSuccess);
    printf("v4 loopback done (%d failures so far)\n", g_fail.load());
  }
  {  // ---- v8: what current NCCL loads (non-blocking connect/accept, tags, grouped irecv)
    ncclNet_v8_t* net = (ncclNet_v8_t*)dlsym(h, "ncclNetPlugin_v8");
    if (!net) { printf("ncclNetPlugin_v8 missing\n"); return 2; }
    CHECK(net->init(nullptr) == ncclSuccess);
    ncclNetProperties_v8_t props;
    CHECK(net->getProperties(0, &props) == ncclSuccess && props.maxRecvs == 1 && props.netDeviceType == NCCL_NET_DEVICE_HOST);
    char handle[NCCL_NET_HANDLE_MAXSIZE];
    void *lcomm = nullptr, *scomm = nullptr, *rcomm = nullptr;
    CHECK(net->listen(0, handle, &lcomm) == ncclSuccess);
    CHECK(net->accept(lcomm, &rcomm, nullptr) == ncclSuccess && rcomm == nullptr);   // nothing connected yet: not ready, no error
    while (!scomm) CHECK(net->connect(0, handle, &scomm, nullptr) == ncclSuccess);
    while (!rcomm) CHECK(net->accept(lcomm, &rcomm, nullptr) == ncclSuccess);
    run_pair(net, scomm, rcomm,
             [&](void* c, void* d, int s, void* mh, void** r) { return net->isend(c, d, s, 7, mh, r); },
             [&](void* c, void* d, int s, void* mh, void** r) {
               int tag = 7;
               return net->irecv(c, 1, &d, &s, &tag, &mh, r);
             });
    CHECK(net->closeSend(scomm) == ncclSuccess && net->closeRecv(rcomm) == ncclSuccess && net->closeListen(lcomm) == ncclSuccess);
  }
  printf("%s (%d failure%s)\n", g_fail ? "FAILED" : "loopback tests passed", g_fail.load(), g_fail == 1 ? "" : "s");
  return g_fail ? 1 : 0;
}
