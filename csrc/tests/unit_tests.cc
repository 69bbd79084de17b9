// C++ unit tests (no GPU): the three utility tests of the reference (reference:
// src/utils.rs:263-314) plus NIC-filter and base64 checks, run against the built .so
// through its exported C API.   usage: unit_tests <path to libnccl-net.so>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <string>

static int g_fail = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) { printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); g_fail++; } \
  } while (0)

template <typename F>
F sym(void* h, const char* name) {
  void* p = dlsym(h, name);
  if (!p) { printf("missing symbol %s\n", name); g_fail++; }
  return reinterpret_cast<F>(p);
}

int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "bagua_net_b200/lib/libnccl-net.so";
  void* h = dlopen(path, RTLD_NOW);
  if (!h) { printf("dlopen(%s): %s\n", path, dlerror()); return 2; }
  auto chunk_count = sym<unsigned long long (*)(unsigned long long, unsigned long long, unsigned long long)>(h, "bnet_chunk_count");
  auto chunk_size = sym<unsigned long long (*)(unsigned long long, unsigned long long, unsigned long long)>(h, "bnet_chunk_size");
  auto parse = sym<int (*)(const char*, char*, char*, char*, int)>(h, "bnet_parse_user_pass_addr");
  auto roundtrip = sym<int (*)(const char*, char*, int)>(h, "bnet_sockaddr_roundtrip");
  auto accepts = sym<int (*)(const char*, const char*)>(h, "bnet_if_filter_accepts");
  auto b64 = sym<int (*)(const char*, char*, int)>(h, "bnet_base64");
  if (g_fail) return 2;

  // test_chunks
  CHECK(chunk_count(1024, 1, 20) == 20);
  CHECK(chunk_count(1024, 1000, 20) == 2);
  for (unsigned long long total : {1ull, 4097ull, 1048577ull, 123456789ull})
    for (unsigned long long n : {1ull, 2ull, 8ull}) {
      unsigned long long cs = chunk_size(total, 65536, n);
      CHECK(cs >= 65536 && cs * n >= total && chunk_count(total, 65536, n) <= n);
    }
  // test_parse
  char u[64], p[64], a[64];
  CHECK(parse("nagle:1984@127.0.0.1:9090", u, p, a, 64) == 0);
  CHECK(!strcmp(u, "nagle") && !strcmp(p, "1984") && !strcmp(a, "127.0.0.1:9090"));
  CHECK(parse("127.0.0.1:9090", u, p, a, 64) == 0 && !u[0] && !p[0] && !strcmp(a, "127.0.0.1:9090"));
  CHECK(parse("", u, p, a, 64) != 0);
  // test_socket_handle
  char out[128];
  CHECK(roundtrip("127.0.0.1:8123", out, 128) == 0 && !strcmp(out, "127.0.0.1:8123"));
  CHECK(roundtrip("[fe80::1]:4242", out, 128) == 0 && !strcmp(out, "[fe80::1]:4242"));
  // NCCL_SOCKET_IFNAME syntax
  CHECK(!accepts("^docker,lo", "lo") && accepts("^docker,lo", "eth0"));
  CHECK(accepts("eth", "eth1") && !accepts("=eth", "eth1") && accepts("=eth1", "eth1"));
  // base64 (Pushgateway basic auth)
  CHECK(b64("user:pass", out, 128) > 0 && !strcmp(out, "dXNlcjpwYXNz"));
  // the NCCL tables must be exported
  for (const char* s : {"ncclNetPlugin_v3", "ncclNetPlugin_v4", "ncclNetPlugin_v6", "ncclNetPlugin_v8"}) CHECK(dlsym(h, s));
  printf("%s (%d failure%s)\n", g_fail ? "FAILED" : "unit tests passed", g_fail, g_fail == 1 ? "" : "s");
  return g_fail ? 1 : 0;
}
