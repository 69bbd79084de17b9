// CUDA-facing helpers for the host engine: availability probes, regMr export /
// import (CUDA IPC for cudaMalloc memory, cuMem POSIX-fd handles for VMM memory —
// NCCL >= 2.19 allocates its buffers that way), pinned host memory, staged copies.
//
// BNET_FAKE_CUDA=1 swaps in a host-memory emulation ("device memory" = named
// POSIX shm segments) so the NVLink transport's control plane — MR exchange,
// FIFO matching, completion flags — is testable on a box without a GPU.
#include "cuda/cuda_iface.h"

#include <cuda_runtime_api.h>
#include <ctype.h>
#include <dlfcn.h>
#include <limits.h>
#include <stdlib.h>
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <map>
#include <mutex>
#include <string>

#include "core/common.h"
#include "cuda/driver_api.h"

namespace bnet {
namespace cuda {

// ------------------------------------------------------------------ driver loader
const char* cu_err(CUresult r) {
  const char* s = nullptr;
  const DriverApi& d = driver();
  if (d.ok && d.GetErrorString && d.GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
  return "unknown CUDA driver error";
}

const DriverApi& driver() {
  static DriverApi api = [] {
    DriverApi a{};
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return a;
    bool all = true;
#define BIND(field, sym)                                   \
  do {                                                     \
    *(void**)(&a.field) = dlsym(h, sym);                   \
    if (!a.field) { all = false; BNET_DEBUG("libcuda: missing %s", sym); } \
  } while (0)
    BIND(GetErrorString, "cuGetErrorString");
    BIND(DeviceGet, "cuDeviceGet");
    BIND(DeviceGetAttribute, "cuDeviceGetAttribute");
    BIND(CtxGetCurrent, "cuCtxGetCurrent");
    BIND(PointerGetAttribute, "cuPointerGetAttribute");
    BIND(MemGetAddressRange, "cuMemGetAddressRange_v2");
    BIND(MemCreate, "cuMemCreate");
    BIND(MemRelease, "cuMemRelease");
    BIND(MemAddressReserve, "cuMemAddressReserve");
    BIND(MemAddressFree, "cuMemAddressFree");
    BIND(MemMap, "cuMemMap");
    BIND(MemUnmap, "cuMemUnmap");
    BIND(MemSetAccess, "cuMemSetAccess");
    BIND(MemGetAllocationGranularity, "cuMemGetAllocationGranularity");
    BIND(MemRetainAllocationHandle, "cuMemRetainAllocationHandle");
    BIND(MemGetAllocationPropertiesFromHandle, "cuMemGetAllocationPropertiesFromHandle");
    BIND(MemExportToShareableHandle, "cuMemExportToShareableHandle");
    BIND(MemImportFromShareableHandle, "cuMemImportFromShareableHandle");
    bool core = all;
    // multicast entry points are optional (older drivers)
    BIND(MulticastCreate, "cuMulticastCreate");
    BIND(MulticastAddDevice, "cuMulticastAddDevice");
    BIND(MulticastBindMem, "cuMulticastBindMem");
    BIND(MulticastUnbind, "cuMulticastUnbind");
    BIND(MulticastGetGranularity, "cuMulticastGetGranularity");
    // stream memory operations (copy-engine transport mode): optional as well
    *(void**)(&a.StreamWriteValue64) = dlsym(h, "cuStreamWriteValue64_v2");
    if (!a.StreamWriteValue64) *(void**)(&a.StreamWriteValue64) = dlsym(h, "cuStreamWriteValue64");
    *(void**)(&a.MemcpyAsync) = dlsym(h, "cuMemcpyAsync");
#undef BIND
    a.ok = core;
    return a;
  }();
  return api;
}

// ------------------------------------------------------------------ fake device memory
namespace {
bool fake_mode() {
  static int f = [] {
    const char* e = getenv("BNET_FAKE_CUDA");
    return (e && atoi(e) == 1) ? 1 : 0;
  }();
  return f == 1;
}
struct FakeSeg {
  std::string name;
  size_t size;
};
std::mutex g_fake_mu;
std::map<uintptr_t, FakeSeg> g_fake;   // base -> segment
}  // namespace

extern "C" __attribute__((visibility("default"))) void* bnet_fake_cuda_alloc(size_t n) {
  static std::atomic<int> ctr{0};
  static std::once_flag once;
  std::call_once(once, [] {
    atexit([] {   // emulated device segments are named shm files: do not leave them behind
      std::lock_guard<std::mutex> lk(g_fake_mu);
      for (auto& kv : g_fake) shm_unlink(kv.second.name.c_str());
    });
  });
  char name[64];
  // pid + counter + random: pids are recycled, and a crashed test may have left its files behind
  snprintf(name, sizeof(name), "/bnet-fake-%d-%d-%08x", (int)getpid(), ctr.fetch_add(1), (unsigned)random_u64());
  int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0) return nullptr;
  if (ftruncate(fd, (off_t)n) != 0) { close(fd); shm_unlink(name); return nullptr; }
  void* p = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) { shm_unlink(name); return nullptr; }
  std::lock_guard<std::mutex> lk(g_fake_mu);
  g_fake[(uintptr_t)p] = FakeSeg{name, n};
  return p;
}

extern "C" __attribute__((visibility("default"))) void bnet_fake_cuda_free(void* p) {
  std::lock_guard<std::mutex> lk(g_fake_mu);
  auto it = g_fake.find((uintptr_t)p);
  if (it == g_fake.end()) return;
  munmap(p, it->second.size);
  shm_unlink(it->second.name.c_str());
  g_fake.erase(it);
}

static bool fake_lookup(const void* p, uintptr_t* base, FakeSeg* seg) {
  std::lock_guard<std::mutex> lk(g_fake_mu);
  auto it = g_fake.upper_bound((uintptr_t)p);
  if (it == g_fake.begin()) return false;
  --it;
  if ((uintptr_t)p >= it->first + it->second.size) return false;
  *base = it->first;
  *seg = it->second;
  return true;
}

// ------------------------------------------------------------------ probes
bool fake() { return fake_mode(); }

int device_count() {
  if (fake_mode()) return 1;
  static int n = [] {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    return c;
  }();
  return n;
}

bool available() { return device_count() > 0; }

int current_device() {
  if (fake_mode()) return 0;
  if (!available()) return -1;
  int d = -1;
  if (cudaGetDevice(&d) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return d;
}

bool pointer_is_device(const void* p, int* dev_out) {
  if (fake_mode()) {
    uintptr_t b;
    FakeSeg s;
    if (dev_out) *dev_out = 0;
    return fake_lookup(p, &b, &s);
  }
  if (!available()) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (dev_out) *dev_out = a.device;
  return a.type == cudaMemoryTypeDevice;
}

std::string device_pci_path(int dev, std::string* busid_out) {
  if (busid_out) busid_out->clear();
  if (fake_mode()) {
    if (busid_out) *busid_out = "0000:00:00.0";
    return "";
  }
  if (!available() || dev < 0 || dev >= device_count()) return "";
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess) {
    cudaGetLastError();
    return "";
  }
  for (char* c = bus; *c; c++) *c = (char)tolower(*c);
  if (busid_out) *busid_out = bus;
  char buf[4096];
  std::string p = std::string("/sys/bus/pci/devices/") + bus;
  if (realpath(p.c_str(), buf)) return buf;
  return "";
}

void* host_device_alias(const void* host_ptr) {
  if (fake_mode()) return const_cast<void*>(host_ptr);   // the emulated "kernel" is a memcpy in this process
  if (!available()) return nullptr;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host_ptr) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (a.type != cudaMemoryTypeHost) return nullptr;       // pageable memory: the GPU cannot read it
  return a.devicePointer;
}

// ------------------------------------------------------------------ export / import
int export_memory(const void* ptr, size_t size, MemExport* out) {
  CallScope cs_("export memory");
  memset(out, 0, sizeof(*out));
  out->pid = (uint64_t)getpid();
  out->fd = -1;
  if (fake_mode()) {
    uintptr_t base;
    FakeSeg seg;
    if (!fake_lookup(ptr, &base, &seg)) return -1;
    out->kind = EXPORT_CUDA_IPC;
    out->dev = 0;
    out->alloc_base = base;
    out->alloc_size = seg.size;
    snprintf((char*)out->ipc, sizeof(out->ipc), "%s", seg.name.c_str());
    return 0;
  }
  const DriverApi& d = driver();
  int dev = 0;
  cudaPointerAttributes pa;
  if (cudaPointerGetAttributes(&pa, ptr) != cudaSuccess || pa.type != cudaMemoryTypeDevice) {
    cudaGetLastError();
    return -1;
  }
  dev = pa.device;
  out->dev = dev;
  CUdeviceptr base = 0;
  size_t asz = 0;
  if (!d.ok || d.MemGetAddressRange(&base, &asz, (CUdeviceptr)ptr) != CUDA_SUCCESS) return -1;
  if ((uintptr_t)ptr + size > (uintptr_t)base + asz) {
    BNET_WARN("regMr: [%p,+%zu) crosses its allocation [%p,+%zu)", ptr, size, (void*)base, asz);
    return -1;
  }
  out->alloc_base = (uint64_t)base;
  out->alloc_size = asz;
  // (1) VMM allocation (cuMemCreate + cuMemMap): export the generic handle as a POSIX fd
  CUmemGenericAllocationHandle h;
  if (d.MemRetainAllocationHandle(&h, (void*)base) == CUDA_SUCCESS) {
    int fd = -1;
    CUresult r = d.MemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    d.MemRelease(h);
    if (r == CUDA_SUCCESS && fd >= 0) {
      out->kind = EXPORT_POSIX_FD;
      out->fd = fd;
      return 0;
    }
    BNET_INFO("regMr: VMM allocation %p is not exportable as a POSIX fd (%s)", (void*)base, cu_err(r));
    return -1;
  }
  // (2) cudaMalloc memory: classic CUDA IPC
  cudaIpcMemHandle_t ipc;
  cudaError_t e = cudaIpcGetMemHandle(&ipc, (void*)base);
  if (e != cudaSuccess) {
    cudaGetLastError();
    BNET_INFO("regMr: cudaIpcGetMemHandle(%p) failed: %s", (void*)base, cudaGetErrorString(e));
    return -1;
  }
  static_assert(sizeof(ipc) <= sizeof(out->ipc), "ipc handle size");
  memcpy(out->ipc, &ipc, sizeof(ipc));
  out->kind = EXPORT_CUDA_IPC;
  return 0;
}

void release_export(MemExport* e) {
  if (e->kind == EXPORT_POSIX_FD && e->fd >= 0) {
    close(e->fd);
    e->fd = -1;
  }
}

// `tag` tells two allocations apart that lived at the same address of the exporter at different times (it frees a
// buffer, maps another one at the same VA, registers it again): the inode behind the POSIX fd, or a hash of the
// CUDA IPC handle.  Without it the cache would hand out the mapping of the OLD physical memory.
struct ImportKey {
  uint64_t pid, base;
  int dev;
  uint64_t tag;
  bool operator<(const ImportKey& o) const {
    if (pid != o.pid) return pid < o.pid;
    if (base != o.base) return base < o.base;
    if (dev != o.dev) return dev < o.dev;
    return tag < o.tag;
  }
};
struct ImportCookie {
  CUmemGenericAllocationHandle h;
  size_t size;
  bool vmm;
  void* base;
  int refs;
  ImportKey key;
};
static std::mutex g_import_mu;
static std::map<ImportKey, ImportCookie*> g_imports;

int import_memory(const MemExport& e, int local_fd, int dev, void** base_out, void** cookie_out) {
  CallScope cs_("import peer memory");
  *base_out = nullptr;
  *cookie_out = nullptr;
  if (fake_mode()) {
    if (e.pid == (uint64_t)getpid()) {
      *base_out = (void*)e.alloc_base;
      return 0;
    }
    int fd = shm_open((const char*)e.ipc, O_RDWR, 0600);
    if (fd < 0) return -1;
    void* p = mmap(nullptr, e.alloc_size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return -1;
    *base_out = p;
    return 0;
  }
  const DriverApi& d = driver();
  if (dev < 0) dev = current_device();
  if (e.pid == (uint64_t)getpid()) {
    // same process (ncclCommInitAll / one process driving several GPUs): no export needed,
    // only access rights for the reading/writing device
    if (dev != e.dev) {
      CUmemGenericAllocationHandle h;
      if (d.ok && d.MemRetainAllocationHandle(&h, (void*)e.alloc_base) == CUDA_SUCCESS) {
        d.MemRelease(h);
        CUmemAccessDesc acc{};
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = dev;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        CUresult r = d.MemSetAccess((CUdeviceptr)e.alloc_base, e.alloc_size, &acc, 1);
        if (r != CUDA_SUCCESS) BNET_INFO("cuMemSetAccess for peer dev %d failed: %s", dev, cu_err(r));
      } else {
        enable_peer_access(dev, e.dev);
      }
    }
    *base_out = (void*)e.alloc_base;
    return 0;
  }
  // One mapping per (exporter, allocation, importing device) and process: many registered
  // buffers usually live in the same allocation, and a CUDA IPC handle must not be opened twice.
  uint64_t tag = 0;
  if (e.kind == EXPORT_CUDA_IPC) {
    tag = fnv1a(e.ipc, sizeof(e.ipc));
  } else if (e.kind == EXPORT_POSIX_FD && local_fd >= 0) {
    struct stat st;
    if (fstat(local_fd, &st) == 0) tag = ((uint64_t)st.st_dev << 32) ^ (uint64_t)st.st_ino;
  }
  ImportKey key{e.pid, e.alloc_base, dev, tag};
  {
    std::lock_guard<std::mutex> lk(g_import_mu);
    auto it = g_imports.find(key);
    if (it != g_imports.end()) {
      it->second->refs++;
      *base_out = it->second->base;
      *cookie_out = it->second;
      return 0;
    }
  }
  if (e.kind == EXPORT_CUDA_IPC) {
    cudaIpcMemHandle_t ipc;
    memcpy(&ipc, e.ipc, sizeof(ipc));
    void* p = nullptr;
    cudaError_t err = cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) {
      cudaGetLastError();
      BNET_INFO("cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(err));
      return -1;
    }
    ImportCookie* c = new ImportCookie{0, (size_t)e.alloc_size, false, p, 1, key};
    std::lock_guard<std::mutex> lk(g_import_mu);
    g_imports[key] = c;
    *base_out = p;
    *cookie_out = c;
    return 0;
  }
  if (e.kind == EXPORT_POSIX_FD) {
    if (!d.ok || local_fd < 0) return -1;
    CUmemGenericAllocationHandle h;
    CUresult r = d.MemImportFromShareableHandle(&h, (void*)(uintptr_t)local_fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    if (r != CUDA_SUCCESS) {
      BNET_INFO("cuMemImportFromShareableHandle failed: %s", cu_err(r));
      return -1;
    }
    CUdeviceptr va = 0;
    size_t sz = (size_t)e.alloc_size;
    CUmemAllocationProp prop{};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = e.dev;
    size_t gran = 2 << 20;
    d.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    if ((r = d.MemAddressReserve(&va, sz, gran, 0, 0)) != CUDA_SUCCESS ||
        (r = d.MemMap(va, sz, 0, h, 0)) != CUDA_SUCCESS) {
      BNET_INFO("cuMemAddressReserve/cuMemMap failed: %s", cu_err(r));
      if (va) d.MemAddressFree(va, sz);
      d.MemRelease(h);
      return -1;
    }
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = dev;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if ((r = d.MemSetAccess(va, sz, &acc, 1)) != CUDA_SUCCESS) {
      BNET_INFO("cuMemSetAccess failed: %s", cu_err(r));
      d.MemUnmap(va, sz);
      d.MemAddressFree(va, sz);
      d.MemRelease(h);
      return -1;
    }
    ImportCookie* c = new ImportCookie{h, sz, true, (void*)va, 1, key};
    std::lock_guard<std::mutex> lk(g_import_mu);
    g_imports[key] = c;
    *base_out = (void*)va;
    *cookie_out = c;
    return 0;
  }
  return -1;
}

void release_import(const MemExport& e, void* base, void* cookie) {
  if (fake_mode()) {
    if (base && e.pid != (uint64_t)getpid()) munmap(base, e.alloc_size);
    return;
  }
  ImportCookie* c = (ImportCookie*)cookie;
  if (!c) return;
  {
    std::lock_guard<std::mutex> lk(g_import_mu);
    if (--c->refs > 0) return;
    g_imports.erase(c->key);
  }
  if (c->vmm) {
    const DriverApi& d = driver();
    d.MemUnmap((CUdeviceptr)base, c->size);
    d.MemAddressFree((CUdeviceptr)base, c->size);
    d.MemRelease(c->h);
  } else {
    cudaIpcCloseMemHandle(base);
  }
  delete c;
}

// ------------------------------------------------------------------ pinned memory
int host_register(void* p, size_t n, void** dev_ptr_out) {
  if (fake_mode() || !available()) {
    if (dev_ptr_out) *dev_ptr_out = p;
    return fake_mode() ? 0 : -1;
  }
  cudaError_t e = cudaHostRegister(p, n, cudaHostRegisterMapped | cudaHostRegisterPortable);
  if (e != cudaSuccess && e != cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError();
    BNET_INFO("cudaHostRegister(%p,%zu) failed: %s", p, n, cudaGetErrorString(e));
    return -1;
  }
  if (e != cudaSuccess) cudaGetLastError();
  void* dp = nullptr;
  if (cudaHostGetDevicePointer(&dp, p, 0) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  if (dev_ptr_out) *dev_ptr_out = dp;
  return 0;
}

int host_unregister(void* p) {
  if (fake_mode() || !available()) return 0;
  cudaError_t e = cudaHostUnregister(p);
  if (e != cudaSuccess) cudaGetLastError();
  return e == cudaSuccess ? 0 : -1;
}

void* host_alloc_mapped(size_t n, void** dev_ptr_out) {
  if (fake_mode() || !available()) {
    void* p = nullptr;
    if (posix_memalign(&p, 4096, n) != 0) return nullptr;
    memset(p, 0, n);
    if (dev_ptr_out) *dev_ptr_out = p;
    return p;
  }
  void* p = nullptr;
  if (cudaHostAlloc(&p, n, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  memset(p, 0, n);
  if (dev_ptr_out) {
    void* dp = nullptr;
    cudaHostGetDevicePointer(&dp, p, 0);
    *dev_ptr_out = dp;
  }
  return p;
}

void host_free_mapped(void* p) {
  if (!p) return;
  if (fake_mode() || !available()) free(p);
  else cudaFreeHost(p);
}

// ------------------------------------------------------------------ staged copies
int memcpy_sync(void* dst, const void* src, size_t n, int dev) {
  CallScope cs_("staged cudaMemcpyAsync+sync");
  if (fake_mode()) {
    memcpy(dst, src, n);
    return 0;
  }
  if (!available()) return -1;
  // A private non-blocking stream: the legacy default stream would serialise behind
  // (and dead-lock with) the NCCL kernel that is waiting for this very transfer.
  thread_local cudaStream_t streams[64] = {nullptr};
  int cur = current_device();
  if (dev >= 0 && dev != cur) {
    if (cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); return -1; }
    cur = dev;
  }
  if (cur < 0 || cur >= 64) return -1;
  if (!streams[cur] && cudaStreamCreateWithFlags(&streams[cur], cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  cudaError_t e = cudaMemcpyAsync(dst, src, n, cudaMemcpyDefault, streams[cur]);
  if (e == cudaSuccess) e = cudaStreamSynchronize(streams[cur]);
  if (e != cudaSuccess) {
    cudaGetLastError();
    BNET_WARN("staged copy of %zu bytes failed: %s", n, cudaGetErrorString(e));
    return -1;
  }
  return 0;
}

int enable_peer_access(int dev, int peer) {
  if (fake_mode() || dev == peer) return 0;
  int can = 0;
  if (cudaDeviceCanAccessPeer(&can, dev, peer) != cudaSuccess || !can) {
    cudaGetLastError();
    return -1;
  }
  int cur = current_device();
  if (cur != dev) cudaSetDevice(dev);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (cur != dev && cur >= 0) cudaSetDevice(cur);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return -1;
  }
  if (e != cudaSuccess) cudaGetLastError();
  return 0;
}

}  // namespace cuda
}  // namespace bnet
