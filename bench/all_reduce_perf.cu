// all_reduce_perf — an nccl-tests-compatible all-reduce sweep (the image ships neither
// nccl-tests nor MPI).  The reference's documented benchmark procedure is
//   all_reduce_perf -b 8 -e 128M -f 2 -g 1          (reference README.md:20,27-44)
// run once without and once with the plugin on LD_LIBRARY_PATH; this clone accepts the same
// flags, forks one process per GPU on this box (instead of mpirun), bootstraps NCCL through a
// unique id in shared memory, times on the device with CUDA events, takes the max over
// ranks and prints size / time / algbw / busbw like nccl-tests.
//
//   build/bench/all_reduce_perf -b 8 -e 128M -f 2 -n 20 -w 5 -N 8 [-d float|half|bfloat16]
// With the plugin:  env $(python -m bagua_net_b200.utils.env) build/bench/all_reduce_perf ...
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <execinfo.h>
#include <signal.h>

#include <atomic>
#include <vector>

// a rank that dies of a signal says where (the plugin is loaded into this process: its frames show up here)
static void crash_handler(int sig) {
  void* frames[64];
  int n = backtrace(frames, 64);
  char msg[96];
  int len = snprintf(msg, sizeof(msg), "\n[all_reduce_perf] pid %d: signal %d, backtrace:\n", (int)getpid(), sig);
  if (write(2, msg, len) < 0) {}
  backtrace_symbols_fd(frames, n, 2);
  _exit(128 + sig);
}

#define CUDACHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(2); } } while (0)
#define NCCLCHECK(x) do { ncclResult_t r = (x); if (r != ncclSuccess) { fprintf(stderr, "NCCL %s:%d %s\n", __FILE__, __LINE__, ncclGetErrorString(r)); exit(3); } } while (0)

struct Shared {
  ncclUniqueId id;
  std::atomic<int> id_ready;
  std::atomic<int> barrier_count[2];
  std::atomic<int> barrier_sense;
  double time_us[64];     // per rank, current size
  int errors[64];
};

static size_t parse_size(const char* s) {
  char* end;
  double v = strtod(s, &end);
  if (*end == 'K' || *end == 'k') v *= 1 << 10;
  else if (*end == 'M' || *end == 'm') v *= 1 << 20;
  else if (*end == 'G' || *end == 'g') v *= 1 << 30;
  return (size_t)v;
}

static void host_barrier(Shared* sh, int n, int* local_sense) {
  *local_sense = !*local_sense;
  int idx = *local_sense;
  if (sh->barrier_count[idx].fetch_add(1) + 1 == n) {
    sh->barrier_count[idx].store(0);
    sh->barrier_sense.store(*local_sense);
  } else {
    while (sh->barrier_sense.load() != *local_sense) usleep(20);
  }
}

template <typename T>
__global__ void fill_kernel(T* p, size_t n, float v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = (T)v;
}
template <typename T>
__global__ void check_kernel(const T* p, size_t n, float expect, int* bad) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if (fabsf((float)p[i] - expect) > 1e-2f * fabsf(expect) + 1e-3f) atomicAdd(bad, 1);
}

static int g_hostid_per_rank = 0;

template <typename T>
static int run_rank(int rank, int nranks, Shared* sh, size_t minb, size_t maxb, double factor, int iters, int warm,
                    ncclDataType_t dt, const char* dtname, int check) {
  if (g_hostid_per_rank) {   // -H: every rank claims to be its own host, so NCCL routes EVERYTHING through the net plugin
    char hid[64];
    snprintf(hid, sizeof(hid), "bnet-vhost-%d", rank);
    setenv("NCCL_HOSTID", hid, 1);
  }
  int ndev = 0;
  CUDACHECK(cudaGetDeviceCount(&ndev));
  CUDACHECK(cudaSetDevice(rank % ndev));
  if (rank == 0) {
    NCCLCHECK(ncclGetUniqueId(&sh->id));
    sh->id_ready.store(1);
  } else {
    while (!sh->id_ready.load()) usleep(100);
  }
  ncclComm_t comm;
  NCCLCHECK(ncclCommInitRank(&comm, nranks, sh->id, rank));
  cudaStream_t st;
  CUDACHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  T *send, *recv;
  CUDACHECK(cudaMalloc(&send, maxb));
  CUDACHECK(cudaMalloc(&recv, maxb));
  int* bad;
  CUDACHECK(cudaMalloc(&bad, sizeof(int)));
  cudaEvent_t e0, e1;
  CUDACHECK(cudaEventCreate(&e0));
  CUDACHECK(cudaEventCreate(&e1));
  int sense = 0;
  if (rank == 0) {
    int ver;
    ncclGetVersion(&ver);
    printf("# nccl %d  nranks %d  dtype %s  (device-timed, max over ranks)\n", ver, nranks, dtname);
    printf("#%11s %12s %8s %10s %8s %8s %7s\n", "size(B)", "count", "type", "time(us)", "algbw", "busbw", "#wrong");
  }
  for (size_t bytes = minb; bytes <= maxb; bytes = (size_t)(bytes * factor) > bytes ? (size_t)(bytes * factor) : bytes + 1) {
    size_t count = bytes / sizeof(T);
    if (count == 0) count = 1;
    fill_kernel<T><<<64, 256, 0, st>>>(send, count, (float)(rank + 1));
    for (int i = 0; i < warm; i++) NCCLCHECK(ncclAllReduce(send, recv, count, dt, ncclSum, comm, st));
    CUDACHECK(cudaStreamSynchronize(st));
    host_barrier(sh, nranks, &sense);
    CUDACHECK(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; i++) NCCLCHECK(ncclAllReduce(send, recv, count, dt, ncclSum, comm, st));
    CUDACHECK(cudaEventRecord(e1, st));
    CUDACHECK(cudaStreamSynchronize(st));
    float ms = 0;
    CUDACHECK(cudaEventElapsedTime(&ms, e0, e1));
    sh->time_us[rank] = ms * 1e3 / iters;
    int nbad = 0;
    if (check) {
      CUDACHECK(cudaMemsetAsync(bad, 0, sizeof(int), st));
      check_kernel<T><<<64, 256, 0, st>>>(recv, count, (float)(nranks * (nranks + 1) / 2), bad);
      CUDACHECK(cudaMemcpyAsync(&nbad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
      CUDACHECK(cudaStreamSynchronize(st));
    }
    sh->errors[rank] = nbad;
    host_barrier(sh, nranks, &sense);
    if (rank == 0) {
      double t = 0;
      int wrong = 0;
      for (int r = 0; r < nranks; r++) {
        if (sh->time_us[r] > t) t = sh->time_us[r];
        wrong += sh->errors[r];
      }
      double algbw = (double)(count * sizeof(T)) / t / 1e3;   // GB/s
      double busbw = algbw * 2.0 * (nranks - 1) / nranks;
      printf("%12zu %12zu %8s %10.2f %8.2f %8.2f %7d\n", count * sizeof(T), count, dtname, t, algbw, busbw, wrong);
      fflush(stdout);
    }
    host_barrier(sh, nranks, &sense);
    if (bytes == maxb) break;
  }
  NCCLCHECK(ncclCommDestroy(comm));
  return 0;
}

int main(int argc, char** argv) {
  size_t minb = 8, maxb = 128 << 20;
  double factor = 2;
  int iters = 20, warm = 5, nranks = 0, check = 1;
  const char* dtype = "float";
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "-b") && i + 1 < argc) minb = parse_size(argv[++i]);
    else if (!strcmp(argv[i], "-e") && i + 1 < argc) maxb = parse_size(argv[++i]);
    else if (!strcmp(argv[i], "-f") && i + 1 < argc) factor = atof(argv[++i]);
    else if (!strcmp(argv[i], "-n") && i + 1 < argc) iters = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-w") && i + 1 < argc) warm = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-N") && i + 1 < argc) nranks = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-g") && i + 1 < argc) ++i;   // accepted for nccl-tests compatibility (1 GPU per process)
    else if (!strcmp(argv[i], "-c") && i + 1 < argc) check = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-d") && i + 1 < argc) dtype = argv[++i];
    else if (!strcmp(argv[i], "-H")) g_hostid_per_rank = 1;
  }
  if (factor <= 1) factor = 2;
  if (nranks <= 0) {
    const char* e = getenv("BNET_NRANKS");
    nranks = e ? atoi(e) : 2;
  }
  if (nranks > 64) nranks = 64;
  Shared* sh = (Shared*)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  memset((void*)sh, 0, sizeof(Shared));
  std::vector<pid_t> kids;
  for (int r = 0; r < nranks; r++) {   // fork BEFORE any CUDA call
    pid_t p = fork();
    if (p == 0) {
      setvbuf(stdout, nullptr, _IOLBF, 0);
      signal(SIGSEGV, crash_handler);
      signal(SIGBUS, crash_handler);
      signal(SIGABRT, crash_handler);
      signal(SIGFPE, crash_handler);
      int rc;
      if (!strcmp(dtype, "half")) rc = run_rank<__half>(r, nranks, sh, minb, maxb, factor, iters, warm, ncclFloat16, "half", check);
      else if (!strcmp(dtype, "bfloat16")) rc = run_rank<__nv_bfloat16>(r, nranks, sh, minb, maxb, factor, iters, warm, ncclBfloat16, "bf16", check);
      else rc = run_rank<float>(r, nranks, sh, minb, maxb, factor, iters, warm, ncclFloat32, "float", check);
      _exit(rc);
    }
    kids.push_back(p);
  }
  int bad = 0;
  for (pid_t p : kids) {
    int stt = 0;
    waitpid(p, &stt, 0);
    if (!WIFEXITED(stt) || WEXITSTATUS(stt) != 0) {
      bad++;
      if (WIFSIGNALED(stt)) fprintf(stderr, "[all_reduce_perf] rank process %d killed by signal %d\n", (int)p, WTERMSIG(stt));
      else fprintf(stderr, "[all_reduce_perf] rank process %d exited with code %d\n", (int)p, WEXITSTATUS(stt));
    }
  }
  return bad ? 1 : 0;
}
