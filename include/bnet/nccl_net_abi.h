/*
 * bnet — NCCL external network plugin ABI, written out by hand.
 *
 * No NCCL plugin headers ship in the image, so every table layout NCCL may
 * dlsym() from us is declared here.  v3/v4 follow the semantics of the
 * reference (reference: cc/v3/nccl_net_v3.h:24-61, cc/v4/nccl_net_v4.h:24-62,
 * cc/nccl_types.h:6-55); v5..v10 follow NCCL's public ext-net contract
 * (SURVEY.md Appendix A).  NCCL 2.27/2.28 only probe v6 and newer, so the
 * tables that matter at run time are v6 and v8+; v3/v4 keep parity with the
 * reference and feed our own loopback harness.
 */
#ifndef BNET_NCCL_NET_ABI_H_
#define BNET_NCCL_NET_ABI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- result codes (nccl.h) ------------------------------------------------ */
typedef enum {
  ncclSuccess = 0,
  ncclUnhandledCudaError = 1,
  ncclSystemError = 2,
  ncclInternalError = 3,
  ncclInvalidArgument = 4,
  ncclInvalidUsage = 5,
  ncclRemoteError = 6,
  ncclInProgress = 7,
  ncclNumResults = 8
} ncclResult_t;

typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3, ncclAvg = 4, ncclNumOps = 5 } ncclRedOp_t;

typedef enum {
  ncclInt8 = 0, ncclChar = 0,
  ncclUint8 = 1,
  ncclInt32 = 2, ncclInt = 2,
  ncclUint32 = 3,
  ncclInt64 = 4,
  ncclUint64 = 5,
  ncclFloat16 = 6, ncclHalf = 6,
  ncclFloat32 = 7, ncclFloat = 7,
  ncclFloat64 = 8, ncclDouble = 8,
  ncclBfloat16 = 9,
  ncclFloat8e4m3 = 10,
  ncclFloat8e5m2 = 11,
  ncclNumTypes = 12
} ncclDataType_t;

/* ---- constants ------------------------------------------------------------- */
#define NCCL_NET_HANDLE_MAXSIZE_V4 64
#define NCCL_NET_HANDLE_MAXSIZE 128

#define NCCL_PTR_HOST 0x1
#define NCCL_PTR_CUDA 0x2
#define NCCL_PTR_DMABUF 0x4

#define NCCL_NET_MAX_REQUESTS_V4 8
#define NCCL_NET_MAX_REQUESTS 32
#define NCCL_NET_MAX_DEVS_PER_NIC 4

/* ---- logging ----------------------------------------------------------------- */
typedef enum {
  NCCL_LOG_NONE = 0, NCCL_LOG_VERSION = 1, NCCL_LOG_WARN = 2, NCCL_LOG_INFO = 3,
  NCCL_LOG_ABORT = 4, NCCL_LOG_TRACE = 5
} ncclDebugLogLevel;

typedef enum {
  NCCL_INIT = 0x1, NCCL_COLL = 0x2, NCCL_P2P = 0x4, NCCL_SHM = 0x8, NCCL_NET = 0x10,
  NCCL_GRAPH = 0x20, NCCL_TUNING = 0x40, NCCL_ENV = 0x80, NCCL_ALLOC = 0x100,
  NCCL_CALL = 0x200, NCCL_PROXY = 0x400, NCCL_NVLS = 0x800, NCCL_BOOTSTRAP = 0x1000,
  NCCL_REG = 0x2000, NCCL_PROFILE = 0x4000, NCCL_RAS = 0x8000, NCCL_ALL = ~0
} ncclDebugLogSubSys;

typedef void (*ncclDebugLogger_t)(ncclDebugLogLevel level, unsigned long flags, const char* file,
                                  int line, const char* fmt, ...);

/* v10 profiler callback: the plugin may report its own events into NCCL's profiler. */
typedef ncclResult_t (*ncclProfilerCallback_t)(void** eHandle, int type, void* pHandle,
                                               int64_t pluginId, void* extData);

/* ---- device-offload descriptor (v7+); we are always a HOST-driven plugin ---- */
typedef enum {
  NCCL_NET_DEVICE_HOST = 0,
  NCCL_NET_DEVICE_UNPACK = 1,
  NCCL_NET_DEVICE_GIN_PROXY = 2,
  NCCL_NET_DEVICE_GIN_GDAKI = 3
} ncclNetDeviceType;
#define NCCL_NET_DEVICE_INVALID_VERSION 0x0

typedef struct {
  ncclNetDeviceType netDeviceType;
  int netDeviceVersion;
  void* handle;
  size_t size;
  int needsProxyProgress;
} ncclNetDeviceHandle_v7_t;
typedef ncclNetDeviceHandle_v7_t ncclNetDeviceHandle_v8_t;
typedef ncclNetDeviceHandle_v7_t ncclNetDeviceHandle_v9_t;
typedef ncclNetDeviceHandle_v7_t ncclNetDeviceHandle_v10_t;

/* =============================== v3 ======================================== */
typedef struct {
  char* name;
  char* pciPath;
  uint64_t guid;
  int ptrSupport;
  int speed;
  int port;
  int maxComms;
} ncclNetProperties_v3_t;
typedef ncclNetProperties_v3_t ncclNetProperties_v4_t;

typedef struct {
  const char* name;
  ncclResult_t (*init)(ncclDebugLogger_t logFunction);
  ncclResult_t (*devices)(int* ndev);
  ncclResult_t (*getProperties)(int dev, ncclNetProperties_v3_t* props);
  ncclResult_t (*listen)(int dev, void* handle, void** listenComm);
  ncclResult_t (*connect)(int dev, void* handle, void** sendComm);
  ncclResult_t (*accept)(void* listenComm, void** recvComm);
  ncclResult_t (*regMr)(void* comm, void* data, int size, int type, void** mhandle);
  ncclResult_t (*deregMr)(void* comm, void* mhandle);
  ncclResult_t (*isend)(void* sendComm, void* data, int size, void* mhandle, void** request);
  ncclResult_t (*irecv)(void* recvComm, void* data, int size, void* mhandle, void** request);
  /* v3: synchronous flush, no request object */
  ncclResult_t (*flush)(void* recvComm, void* data, int size, void* mhandle);
  ncclResult_t (*test)(void* request, int* done, int* size);
  ncclResult_t (*closeSend)(void* sendComm);
  ncclResult_t (*closeRecv)(void* recvComm);
  ncclResult_t (*closeListen)(void* listenComm);
} ncclNet_v3_t;

/* =============================== v4 ======================================== */
typedef struct {
  const char* name;
  ncclResult_t (*init)(ncclDebugLogger_t logFunction);
  ncclResult_t (*devices)(int* ndev);
  ncclResult_t (*getProperties)(int dev, ncclNetProperties_v4_t* props);
  ncclResult_t (*listen)(int dev, void* handle, void** listenComm);
  ncclResult_t (*connect)(int dev, void* handle, void** sendComm);
  ncclResult_t (*accept)(void* listenComm, void** recvComm);
  ncclResult_t (*regMr)(void* comm, void* data, int size, int type, void** mhandle);
  ncclResult_t (*deregMr)(void* comm, void* mhandle);
  ncclResult_t (*isend)(void* sendComm, void* data, int size, void* mhandle, void** request);
  ncclResult_t (*irecv)(void* recvComm, void* data, int size, void* mhandle, void** request);
  /* v4: asynchronous flush, completion through test() */
  ncclResult_t (*iflush)(void* recvComm, void* data, int size, void* mhandle, void** request);
  ncclResult_t (*test)(void* request, int* done, int* size);
  ncclResult_t (*closeSend)(void* sendComm);
  ncclResult_t (*closeRecv)(void* recvComm);
  ncclResult_t (*closeListen)(void* listenComm);
} ncclNet_v4_t;

/* Collective-offload table of the v4 era.  The reference only carries the
 * declaration (cc/v4/nccl_net_v4.h:64-101) and exports no ncclCollNetPlugin symbol;
 * ours is exported too (csrc/plugin/collnet.cc), next to the v6..v10 tables at the
 * end of this header that the installed NCCLs actually probe. */
typedef struct {
  const char* name;
  ncclResult_t (*init)(ncclDebugLogger_t logFunction);
  ncclResult_t (*devices)(int* ndev);
  ncclResult_t (*getProperties)(int dev, ncclNetProperties_v4_t* props);
  ncclResult_t (*listen)(int dev, void* handle, void** listenComm);
  ncclResult_t (*connect)(void* handles[], int nranks, int rank, void* listenComm, void** collComm);
  ncclResult_t (*reduceSupport)(ncclDataType_t dataType, ncclRedOp_t redOp, int* supported);
  ncclResult_t (*regMr)(void* collComm, void* data, int size, int type, void** mhandle);
  ncclResult_t (*deregMr)(void* collComm, void* mhandle);
  ncclResult_t (*iallreduce)(void* collComm, void* sendData, void* recvData, int count,
                             ncclDataType_t dataType, ncclRedOp_t redOp, void* sendMhandle,
                             void* recvMhandle, void** request);
  ncclResult_t (*iflush)(void* collComm, void* data, int size, void* mhandle, void** request);
  ncclResult_t (*test)(void* request, int* done, int* size);
  ncclResult_t (*closeColl)(void* collComm);
  ncclResult_t (*closeListen)(void* listenComm);
} ncclCollNet_v4_t;

/* =============================== v5 / v6 ==================================== */
typedef struct {
  char* name;
  char* pciPath;
  uint64_t guid;
  int ptrSupport;
  int speed;
  int port;
  float latency;
  int maxComms;
  int maxRecvs;
} ncclNetProperties_v6_t;
typedef ncclNetProperties_v6_t ncclNetProperties_v5_t;

typedef struct {
  const char* name;
  ncclResult_t (*init)(ncclDebugLogger_t logFunction);
  ncclResult_t (*devices)(int* ndev);
  ncclResult_t (*getProperties)(int dev, ncclNetProperties_v5_t* props);
  ncclResult_t (*listen)(int dev, void* handle, void** listenComm);
  ncclResult_t (*connect)(int dev, void* handle, void** sendComm);
  ncclResult_t (*accept)(void* listenComm, void** recvComm);
  ncclResult_t (*regMr)(void* comm, void* data, int size, int type, void** mhandle);
  ncclResult_t (*deregMr)(void* comm, void* mhandle);
  ncclResult_t (*isend)(void* sendComm, void* data, int size, int tag, void* mhandle, void** request);
  ncclResult_t (*irecv)(void* recvComm, int n, void** data, int* sizes, int* tags, void** mhandles,
                        void** request);
  ncclResult_t (*iflush)(void* recvComm, int n, void** data, int* sizes, void** mhandles,
                         void** request);
  ncclResult_t (*test)(void* request, int* done, int* sizes);
  ncclResult_t (*closeSend)(void* sendComm);
  ncclResult_t (*closeRecv)(void* recvComm);
  ncclResult_t (*closeListen)(void* listenComm);
} ncclNet_v5_t;

typedef struct {
  const char* name;
  ncclResult_t (*init)(ncclDebugLogger_t logFunction);
  ncclResult_t (*devices)(int* ndev);
  ncclResult_t (*getProperties)(int dev, ncclNetProperties_v6_t* props);
  ncclResult_t (*listen)(int dev, void* handle, void** listenComm);
  ncclResult_t (*connect)(int dev, void* handle, void** sendComm);
  ncclResult_t (*accept)(void* listenComm, void** recvComm);
  ncclResult_t (*regMr)(void* comm, void* data, int size, int type, void** mhandle);
  ncclResult_t (*regMrDmaBuf)(void* comm, void* data, size_t size, int type, uint64_t offset, int fd,
                              void** mhandle);
  ncclResult_t (*deregMr)(void* comm, void* mhandle);
  ncclResult_t (*isend)(void* sendComm, void* data, int size, int tag, void* mhandle, void** request);
  ncclResult_t (*irecv)(void* recvComm, int n, void** data, int* sizes, int* tags, void** mhandles,
                        void** request);
  ncclResult_t (*iflush)(void* recvComm, int n, void** data, int* sizes, void** mhandles,
                         void** request);
  ncclResult_t (*test)(void* request, int* done, int* sizes);
  ncclResult_t (*closeSend)(void* sendComm);
  ncclResult_t (*closeRecv)(void* recvComm);
  ncclResult_t (*closeListen)(void* listenComm);
} ncclNet_v6_t;

/* =============================== v7 ======================================== */
typedef struct {
  char* name;
  char* pciPath;
  uint64_t guid;
  int ptrSupport;
  int speed;
  int port;
  float latency;
  int maxComms;
  int maxRecvs;
  ncclNetDeviceType netDeviceType;
  int netDeviceVersion;
} ncclNetProperties_v7_t;

typedef struct {
  const char* name;
  ncclResult_t (*init)(ncclDebugLogger_t logFunction);
  ncclResult_t (*devices)(int* ndev);
  ncclResult_t (*getProperties)(int dev, ncclNetProperties_v7_t* props);
  ncclResult_t (*listen)(int dev, void* handle, void** listenComm);
  ncclResult_t (*connect)(int dev, void* handle, void** sendComm, ncclNetDeviceHandle_v7_t** sendDevComm);
  ncclResult_t (*accept)(void* listenComm, void** recvComm, ncclNetDeviceHandle_v7_t** recvDevComm);
  ncclResult_t (*regMr)(void* comm, void* data, int size, int type, void** mhandle);
  ncclResult_t (*regMrDmaBuf)(void* comm, void* data, size_t size, int type, uint64_t offset, int fd,
                              void** mhandle);
  ncclResult_t (*deregMr)(void* comm, void* mhandle);
  ncclResult_t (*isend)(void* sendComm, void* data, int size, int tag, void* mhandle, void** request);
  ncclResult_t (*irecv)(void* recvComm, int n, void** data, int* sizes, int* tags, void** mhandles,
                        void** request);
  ncclResult_t (*iflush)(void* recvComm, int n, void** data, int* sizes, void** mhandles,
                         void** request);
  ncclResult_t (*test)(void* request, int* done, int* sizes);
  ncclResult_t (*closeSend)(void* sendComm);
  ncclResult_t (*closeRecv)(void* recvComm);
  ncclResult_t (*closeListen)(void* listenComm);
  ncclResult_t (*getDeviceMr)(void* comm, void* mhandle, void** dptr_mhandle);
  ncclResult_t (*irecvConsumed)(void* recvComm, int n, void* request);
} ncclNet_v7_t;

/* =============================== v8 ======================================== */
typedef struct {
  char* name;
  char* pciPath;
  uint64_t guid;
  int ptrSupport;
  int regIsGlobal;
  int speed;
  int port;
  float latency;
  int maxComms;
  int maxRecvs;
  ncclNetDeviceType netDeviceType;
  int netDeviceVersion;
} ncclNetProperties_v8_t;

typedef struct {
  const char* name;
  ncclResult_t (*init)(ncclDebugLogger_t logFunction);
  ncclResult_t (*devices)(int* ndev);
  ncclResult_t (*getProperties)(int dev, ncclNetProperties_v8_t* props);
  ncclResult_t (*listen)(int dev, void* handle, void** listenComm);
  ncclResult_t (*connect)(int dev, void* handle, void** sendComm, ncclNetDeviceHandle_v8_t** sendDevComm);
  ncclResult_t (*accept)(void* listenComm, void** recvComm, ncclNetDeviceHandle_v8_t** recvDevComm);
  ncclResult_t (*regMr)(void* comm, void* data, size_t size, int type, void** mhandle);
  ncclResult_t (*regMrDmaBuf)(void* comm, void* data, size_t size, int type, uint64_t offset, int fd,
                              void** mhandle);
  ncclResult_t (*deregMr)(void* comm, void* mhandle);
  ncclResult_t (*isend)(void* sendComm, void* data, int size, int tag, void* mhandle, void** request);
  ncclResult_t (*irecv)(void* recvComm, int n, void** data, int* sizes, int* tags, void** mhandles,
                        void** request);
  ncclResult_t (*iflush)(void* recvComm, int n, void** data, int* sizes, void** mhandles,
                         void** request);
  ncclResult_t (*test)(void* request, int* done, int* sizes);
  ncclResult_t (*closeSend)(void* sendComm);
  ncclResult_t (*closeRecv)(void* recvComm);
  ncclResult_t (*closeListen)(void* listenComm);
  ncclResult_t (*getDeviceMr)(void* comm, void* mhandle, void** dptr_mhandle);
  ncclResult_t (*irecvConsumed)(void* recvComm, int n, void* request);
} ncclNet_v8_t;

/* =============================== v9 ======================================== */
typedef struct {
  int ndevs;
  int devs[NCCL_NET_MAX_DEVS_PER_NIC];
} ncclNetVDeviceProps_v9_t;

typedef struct {
  char* name;
  char* pciPath;
  uint64_t guid;
  int ptrSupport;
  int regIsGlobal;
  int forceFlush;
  int speed;
  int port;
  float latency;
  int maxComms;
  int maxRecvs;
  ncclNetDeviceType netDeviceType;
  int netDeviceVersion;
  ncclNetVDeviceProps_v9_t vProps;
  size_t maxP2pBytes;
  size_t maxCollBytes;
} ncclNetProperties_v9_t;

typedef struct {
  const char* name;
  ncclResult_t (*init)(ncclDebugLogger_t logFunction);
  ncclResult_t (*devices)(int* ndev);
  ncclResult_t (*getProperties)(int dev, ncclNetProperties_v9_t* props);
  ncclResult_t (*listen)(int dev, void* handle, void** listenComm);
  ncclResult_t (*connect)(int dev, void* handle, void** sendComm, ncclNetDeviceHandle_v9_t** sendDevComm);
  ncclResult_t (*accept)(void* listenComm, void** recvComm, ncclNetDeviceHandle_v9_t** recvDevComm);
  ncclResult_t (*regMr)(void* comm, void* data, size_t size, int type, void** mhandle);
  ncclResult_t (*regMrDmaBuf)(void* comm, void* data, size_t size, int type, uint64_t offset, int fd,
                              void** mhandle);
  ncclResult_t (*deregMr)(void* comm, void* mhandle);
  ncclResult_t (*isend)(void* sendComm, void* data, size_t size, int tag, void* mhandle, void** request);
  ncclResult_t (*irecv)(void* recvComm, int n, void** data, size_t* sizes, int* tags, void** mhandles,
                        void** request);
  ncclResult_t (*iflush)(void* recvComm, int n, void** data, int* sizes, void** mhandles,
                         void** request);
  ncclResult_t (*test)(void* request, int* done, int* sizes);
  ncclResult_t (*closeSend)(void* sendComm);
  ncclResult_t (*closeRecv)(void* recvComm);
  ncclResult_t (*closeListen)(void* listenComm);
  ncclResult_t (*getDeviceMr)(void* comm, void* mhandle, void** dptr_mhandle);
  ncclResult_t (*irecvConsumed)(void* recvComm, int n, void* request);
  ncclResult_t (*makeVDevice)(int* d, ncclNetVDeviceProps_v9_t* props);
} ncclNet_v9_t;

/* =============================== v10 ======================================= */
typedef ncclNetVDeviceProps_v9_t ncclNetVDeviceProps_v10_t;
typedef ncclNetProperties_v9_t ncclNetProperties_v10_t;
typedef struct {
  int trafficClass;
} ncclNetCommConfig_v10_t;

typedef struct {
  const char* name;
  ncclResult_t (*init)(ncclDebugLogger_t logFunction, ncclProfilerCallback_t profFunction);
  ncclResult_t (*devices)(int* ndev);
  ncclResult_t (*getProperties)(int dev, ncclNetProperties_v10_t* props);
  ncclResult_t (*listen)(int dev, void* handle, void** listenComm);
  ncclResult_t (*connect)(int dev, ncclNetCommConfig_v10_t* config, void* handle, void** sendComm,
                          ncclNetDeviceHandle_v10_t** sendDevComm);
  ncclResult_t (*accept)(void* listenComm, void** recvComm, ncclNetDeviceHandle_v10_t** recvDevComm);
  ncclResult_t (*regMr)(void* comm, void* data, size_t size, int type, void** mhandle);
  ncclResult_t (*regMrDmaBuf)(void* comm, void* data, size_t size, int type, uint64_t offset, int fd,
                              void** mhandle);
  ncclResult_t (*deregMr)(void* comm, void* mhandle);
  ncclResult_t (*isend)(void* sendComm, void* data, size_t size, int tag, void* mhandle, void* phandle,
                        void** request);
  ncclResult_t (*irecv)(void* recvComm, int n, void** data, size_t* sizes, int* tags, void** mhandles,
                        void** phandles, void** request);
  ncclResult_t (*iflush)(void* recvComm, int n, void** data, int* sizes, void** mhandles,
                         void** request);
  ncclResult_t (*test)(void* request, int* done, int* sizes);
  ncclResult_t (*closeSend)(void* sendComm);
  ncclResult_t (*closeRecv)(void* recvComm);
  ncclResult_t (*closeListen)(void* listenComm);
  ncclResult_t (*getDeviceMr)(void* comm, void* mhandle, void** dptr_mhandle);
  ncclResult_t (*irecvConsumed)(void* recvComm, int n, void* request);
  ncclResult_t (*makeVDevice)(int* d, ncclNetVDeviceProps_v10_t* props);
} ncclNet_v10_t;

/* ====================== collective offload (CollNet), v6 .. v10 =================
 * What NCCL dlsym()s as ncclCollNetPlugin_vN next to the net table of the same library (NCCL 2.27 / 2.28 probe
 * v10, v9, v8, v7, v6).  The reference carries only the v4 declaration above and exports nothing; here the tables
 * are implemented by csrc/plugin/collnet.cc on top of the two-shot all-reduce of csrc/coll/transport_mesh.cc.
 * Layouts written from NCCL's public ext-net contract (from memory, verify by loading). */
typedef struct {
  const char* name;
  ncclResult_t (*init)(ncclDebugLogger_t logFunction);
  ncclResult_t (*devices)(int* ndev);
  ncclResult_t (*getProperties)(int dev, ncclNetProperties_v6_t* props);
  ncclResult_t (*listen)(int dev, void* handle, void** listenComm);
  ncclResult_t (*connect)(void* handles[], int nranks, int rank, void* listenComm, void** collComm);
  ncclResult_t (*reduceSupport)(ncclDataType_t dataType, ncclRedOp_t redOp, int* supported);
  ncclResult_t (*regMr)(void* collComm, void* data, int size, int type, void** mhandle);
  ncclResult_t (*regMrDmaBuf)(void* collComm, void* data, size_t size, int type, uint64_t offset, int fd,
                              void** mhandle);
  ncclResult_t (*deregMr)(void* collComm, void* mhandle);
  ncclResult_t (*iallreduce)(void* collComm, void* sendData, void* recvData, int count,
                             ncclDataType_t dataType, ncclRedOp_t redOp, void* sendMhandle,
                             void* recvMhandle, void** request);
  ncclResult_t (*iflush)(void* collComm, void* data, int size, void* mhandle, void** request);
  ncclResult_t (*test)(void* request, int* done, int* size);
  ncclResult_t (*closeColl)(void* collComm);
  ncclResult_t (*closeListen)(void* listenComm);
} ncclCollNet_v6_t;

typedef struct {
  const char* name;
  ncclResult_t (*init)(ncclDebugLogger_t logFunction);
  ncclResult_t (*devices)(int* ndev);
  ncclResult_t (*getProperties)(int dev, ncclNetProperties_v7_t* props);
  ncclResult_t (*listen)(int dev, void* handle, void** listenComm);
  ncclResult_t (*connect)(void* handles[], int nranks, int rank, void* listenComm, void** collComm);
  ncclResult_t (*reduceSupport)(ncclDataType_t dataType, ncclRedOp_t redOp, int* supported);
  ncclResult_t (*regMr)(void* collComm, void* data, int size, int type, void** mhandle);
  ncclResult_t (*regMrDmaBuf)(void* collComm, void* data, size_t size, int type, uint64_t offset, int fd,
                              void** mhandle);
  ncclResult_t (*deregMr)(void* collComm, void* mhandle);
  ncclResult_t (*iallreduce)(void* collComm, void* sendData, void* recvData, int count,
                             ncclDataType_t dataType, ncclRedOp_t redOp, void* sendMhandle,
                             void* recvMhandle, void** request);
  ncclResult_t (*iflush)(void* collComm, void* data, int size, void* mhandle, void** request);
  ncclResult_t (*test)(void* request, int* done, int* size);
  ncclResult_t (*closeColl)(void* collComm);
  ncclResult_t (*closeListen)(void* listenComm);
} ncclCollNet_v7_t;

typedef struct {
  void* mhandle;
  void* address;
  uint32_t size;
} ncclNetSGE_v8_t;

typedef struct {
  const char* name;
  ncclResult_t (*init)(ncclDebugLogger_t logFunction);
  ncclResult_t (*devices)(int* ndev);
  ncclResult_t (*getProperties)(int dev, ncclNetProperties_v8_t* props);
  ncclResult_t (*listen)(int dev, void* handle, void** listenComm);
  ncclResult_t (*connect)(void* handles[], int nranks, int rank, void* listenComm, void** collComm);
  ncclResult_t (*reduceSupport)(ncclDataType_t dataType, ncclRedOp_t redOp, int* supported);
  ncclResult_t (*regMr)(void* collComm, void* data, size_t size, int type, void** mhandle);
  ncclResult_t (*regMrDmaBuf)(void* collComm, void* data, size_t size, int type, uint64_t offset, int fd,
                              void** mhandle);
  ncclResult_t (*deregMr)(void* collComm, void* mhandle);
  ncclResult_t (*iallreduce)(void* collComm, void* sendData, void* recvData, int count,
                             ncclDataType_t dataType, ncclRedOp_t redOp, void* sendMhandle,
                             void* recvMhandle, void** request);
  ncclResult_t (*iallgather)(void* collComm, void* sendData, int nRecvParts, ncclNetSGE_v8_t* recvParts,
                             size_t bytesPerRank, size_t windowOffset, size_t windowBytes,
                             void* sendMhandle, void** request);
  ncclResult_t (*ireducescatter)(void* collComm, int nSendParts, ncclNetSGE_v8_t* sendParts, void* recvData,
                                 size_t bytesPerRank, size_t windowOffset, size_t windowBytes,
                                 ncclDataType_t dataType, ncclRedOp_t redOp, void* recvMhandle,
                                 void** request);
  ncclResult_t (*iflush)(void* collComm, void* data, int size, void* mhandle, void** request);
  ncclResult_t (*test)(void* request, int* done, int* size);
  ncclResult_t (*closeColl)(void* collComm);
  ncclResult_t (*closeListen)(void* listenComm);
} ncclCollNet_v8_t;

typedef struct {
  void* mhandle;
  void* address;
  size_t size;
} ncclNetSGE_v9_t;

typedef struct {
  const char* name;
  ncclResult_t (*init)(ncclDebugLogger_t logFunction);
  ncclResult_t (*devices)(int* ndev);
  ncclResult_t (*getProperties)(int dev, ncclNetProperties_v9_t* props);
  ncclResult_t (*listen)(int dev, void* handle, void** listenComm);
  ncclResult_t (*connect)(void* handles[], int nranks, int rank, void* listenComm, void** collComm);
  ncclResult_t (*reduceSupport)(ncclDataType_t dataType, ncclRedOp_t redOp, int* supported);
  ncclResult_t (*regMr)(void* collComm, void* data, size_t size, int type, void** mhandle);
  ncclResult_t (*regMrDmaBuf)(void* collComm, void* data, size_t size, int type, uint64_t offset, int fd,
                              void** mhandle);
  ncclResult_t (*deregMr)(void* collComm, void* mhandle);
  ncclResult_t (*iallreduce)(void* collComm, void* sendData, void* recvData, size_t count,
                             ncclDataType_t dataType, ncclRedOp_t redOp, void* sendMhandle,
                             void* recvMhandle, void** request);
  ncclResult_t (*iallgather)(void* collComm, void* sendData, int nRecvParts, ncclNetSGE_v9_t* recvParts,
                             size_t bytesPerRank, size_t windowOffset, size_t windowBytes,
                             void* sendMhandle, void** request);
  ncclResult_t (*ireducescatter)(void* collComm, int nSendParts, ncclNetSGE_v9_t* sendParts, void* recvData,
                                 size_t bytesPerRank, size_t windowOffset, size_t windowBytes,
                                 ncclDataType_t dataType, ncclRedOp_t redOp, void* recvMhandle,
                                 void** request);
  ncclResult_t (*iflush)(void* collComm, void* data, int size, void* mhandle, void** request);
  ncclResult_t (*test)(void* request, int* done, int* size);
  ncclResult_t (*closeColl)(void* collComm);
  ncclResult_t (*closeListen)(void* listenComm);
  ncclResult_t (*makeVDevice)(int* d, ncclNetVDeviceProps_v9_t* props);
} ncclCollNet_v9_t;

typedef ncclNetSGE_v9_t ncclNetSGE_v10_t;
typedef ncclCollNet_v9_t ncclCollNet_v10_t;   /* same entry points; the properties type is shared with v9 already */

#ifdef __cplusplus
}
#endif
#endif /* BNET_NCCL_NET_ABI_H_ */
