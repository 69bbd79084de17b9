// Lazily resolved CUDA driver API (libcuda.so.1 is dlopen()ed on first use so the
// plugin still loads — and the TCP/shared-memory paths still work — on a box
// without a driver).  Only the entry points bnet needs are bound.
#pragma once
#include <cuda.h>

namespace bnet {
namespace cuda {

struct DriverApi {
  bool ok = false;
  CUresult (*GetErrorString)(CUresult, const char**);
  CUresult (*DeviceGet)(CUdevice*, int);
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice);
  CUresult (*CtxGetCurrent)(CUcontext*);
  CUresult (*PointerGetAttribute)(void*, CUpointer_attribute, CUdeviceptr);
  CUresult (*MemGetAddressRange)(CUdeviceptr*, size_t*, CUdeviceptr);
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  CUresult (*MemRelease)(CUmemGenericAllocationHandle);
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*MemAddressFree)(CUdeviceptr, size_t);
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*MemUnmap)(CUdeviceptr, size_t);
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  CUresult (*MemRetainAllocationHandle)(CUmemGenericAllocationHandle*, void*);
  CUresult (*MemGetAllocationPropertiesFromHandle)(CUmemAllocationProp*, CUmemGenericAllocationHandle);
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice);
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long);
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t);
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
  CUresult (*StreamWriteValue64)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);   // optional
  CUresult (*MemcpyAsync)(CUdeviceptr, CUdeviceptr, size_t, CUstream);                // optional
};

const DriverApi& driver();            // .ok == false when libcuda is missing
const char* cu_err(CUresult r);

}  // namespace cuda
}  // namespace bnet
