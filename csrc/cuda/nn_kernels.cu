// Fused memory-bound layers of the benchmark models (channels_last / NHWC), sm_100a.
//
// profiles/step_profile_torch.txt: in an eager VGG16 bf16 step on B200 the tensor-core convolutions
// are ~30% of the time; bias-add, ReLU, max-pool, their backward passes and the bias-gradient
// reductions — all HBM-bound, one full pass over the activations each — are > 50%.  These kernels
// fuse every such chain behind a convolution into ONE pass:
//   forward : z = conv(x,w)  ->  y = relu(z + b)                      (in place, 1 read + 1 write)
//             z = conv(x,w)  ->  p = maxpool2x2(relu(z + b)), idx      (y is never materialised)
//   backward: gz = gy * (y > 0)  and  gb = sum(gz)                     (1 pass instead of 2.3)
//             gz = scatter(gp, idx) and gb = sum(gp | valid)           (1 pass instead of 3+,
//                                                                      uint8 indices instead of int64)
// All accesses are 16-byte vectors along the contiguous channel dimension.  The per-thread walks live in
// nn_body.cuh so that `make test` can run them over an emulated grid on the CPU (csrc/tests/nn_emu_test.cc).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "cuda/nn_body.cuh"

#define BNET_API extern "C" __attribute__((visibility("default")))

namespace bnet {
namespace nn {

// ---- y = relu(z + b), in place ------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bias_relu_kernel(T* __restrict__ z, const T* __restrict__ bias, size_t nvec, int cvec) {
  bias_relu_thread<T>(z, bias, nvec, cvec, (size_t)blockIdx.x * blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x);
}

// block-wide reduction of per-thread channel-group sums, then one atomic per channel per block
template <int V>
__device__ __forceinline__ void reduce_bias_grad(float* acc, float* __restrict__ gb, int cvec, float* smem) {
  const int rows = blockDim.x / cvec;           // threads that share a channel group
  const int grp = threadIdx.x % cvec, trow = threadIdx.x / cvec;
#pragma unroll
  for (int k = 0; k < V; k++) smem[(trow * cvec + grp) * V + k] = acc[k];
  __syncthreads();
  if (trow == 0) {
    float s[V];
#pragma unroll
    for (int k = 0; k < V; k++) s[k] = 0.f;
    for (int r = 0; r < rows; r++)
#pragma unroll
      for (int k = 0; k < V; k++) s[k] += smem[(r * cvec + grp) * V + k];
#pragma unroll
    for (int k = 0; k < V; k++) atomicAdd(gb + grp * V + k, s[k]);
  }
}

// ---- gz = gy * (y > 0) ; gb += column sums of gz ---------------------------------------------------
// blockDim.x is a multiple of cvec; the per-thread walk (4 rows = 8 x 16-byte loads in flight) is nn_body.cuh.
template <typename T>
__global__ void __launch_bounds__(256) relu_bwd_bias_grad_kernel(const T* __restrict__ gy, const T* __restrict__ y,
                                                                 T* __restrict__ gz, float* __restrict__ gb, size_t rows,
                                                                 int cvec) {
  constexpr int V = Vec<T>::N;
  extern __shared__ float smem[];
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; k++) acc[k] = 0.f;
  relu_bwd_thread<T, 4>(gy, y, gz, rows, cvec, blockDim.x / cvec, threadIdx.x % cvec, threadIdx.x / cvec, blockIdx.x,
                        gridDim.x, acc);
  reduce_bias_grad<V>(acc, gb, cvec, smem);
}

// ---- p = maxpool2x2(relu(z + b)) with a 1-byte index per element ---------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bias_relu_pool_fwd_kernel(const T* __restrict__ z, const T* __restrict__ bias,
                                                                 T* __restrict__ p, uint8_t* __restrict__ idx, int N, int H,
                                                                 int W, int cvec) {
  bias_relu_pool_fwd_thread<T>(z, bias, p, idx, N, H, W, cvec, (size_t)blockIdx.x * blockDim.x + threadIdx.x,
                               (size_t)gridDim.x * blockDim.x);
}

// ---- gz = scatter(gp through idx) ; gb += sum of routed gradients -----------------------------------
template <typename T>
__global__ void __launch_bounds__(256) pool_relu_bwd_bias_grad_kernel(const T* __restrict__ gp, const uint8_t* __restrict__ idx,
                                                                      T* __restrict__ gz, float* __restrict__ gb, int N, int H,
                                                                      int W, int cvec) {
  constexpr int V = Vec<T>::N;
  extern __shared__ float smem[];
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; k++) acc[k] = 0.f;
  pool_relu_bwd_thread<T, 4>(gp, idx, gz, N, H, W, cvec, blockDim.x / cvec, threadIdx.x % cvec, threadIdx.x / cvec,
                             blockIdx.x, gridDim.x, acc);
  reduce_bias_grad<V>(acc, gb, cvec, smem);
}

}  // namespace nn
}  // namespace bnet

using namespace bnet::nn;

// dtype: 0 = f32, 1 = bf16 (same codes as bnet_coll.h).  All tensors NHWC-contiguous, C % vec == 0.
// Return 1 (kernels launched) or <0.
BNET_API int bnet_nn_bias_relu(void* z, const void* bias, long long rows, int C, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int V = dtype == 1 ? 8 : 4;
  if (C % V) return -1;
  size_t nvec = (size_t)rows * C / V;
  int grid = (int)((nvec + 255) / 256 < 148 * 16 ? (nvec + 255) / 256 : 148 * 16);
  if (grid < 1) grid = 1;
  if (dtype == 1) bias_relu_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16*)z, (const __nv_bfloat16*)bias, nvec, C / V);
  else bias_relu_kernel<float><<<grid, 256, 0, st>>>((float*)z, (const float*)bias, nvec, C / V);
  return cudaGetLastError() == cudaSuccess ? 1 : -2;
}

BNET_API int bnet_nn_relu_bwd_bias_grad(const void* gy, const void* y, void* gz, float* gb, long long rows, int C, int dtype,
                                        void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int V = dtype == 1 ? 8 : 4;
  if (C % V || C / V > 256) return -1;
  const int cvec = C / V;
  const int threads = (256 / cvec) * cvec;
  const int rpb = threads / cvec;
  long long want = (rows + rpb - 1) / rpb;
  int grid = (int)(want < 148 * 8 ? want : 148 * 8);
  if (grid < 1) grid = 1;
  size_t smem = (size_t)threads * V * sizeof(float);
  if (dtype == 1)
    relu_bwd_bias_grad_kernel<__nv_bfloat16><<<grid, threads, smem, st>>>((const __nv_bfloat16*)gy, (const __nv_bfloat16*)y,
                                                                       (__nv_bfloat16*)gz, gb, (size_t)rows, cvec);
  else
    relu_bwd_bias_grad_kernel<float><<<grid, threads, smem, st>>>((const float*)gy, (const float*)y, (float*)gz, gb, (size_t)rows, cvec);
  return cudaGetLastError() == cudaSuccess ? 1 : -2;
}

BNET_API int bnet_nn_bias_relu_pool_fwd(const void* z, const void* bias, void* p, void* idx, int N, int H, int W, int C, int dtype,
                                        void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int V = dtype == 1 ? 8 : 4;
  if (C % V || (H & 1) || (W & 1)) return -1;
  size_t total = (size_t)N * (H / 2) * (W / 2) * (C / V);
  int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (grid < 1) grid = 1;
  if (dtype == 1)
    bias_relu_pool_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)z, (const __nv_bfloat16*)bias,
                                                                  (__nv_bfloat16*)p, (uint8_t*)idx, N, H, W, C / V);
  else
    bias_relu_pool_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)z, (const float*)bias, (float*)p, (uint8_t*)idx, N, H, W, C / V);
  return cudaGetLastError() == cudaSuccess ? 1 : -2;
}

BNET_API int bnet_nn_pool_relu_bwd_bias_grad(const void* gp, const void* idx, void* gz, float* gb, int N, int H, int W, int C,
                                             int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int V = dtype == 1 ? 8 : 4;
  if (C % V || C / V > 256 || (H & 1) || (W & 1)) return -1;
  const int cvec = C / V;
  const int threads = (256 / cvec) * cvec;
  const int rpb = threads / cvec;
  long long rows = (long long)N * (H / 2) * (W / 2);
  long long want = (rows + rpb - 1) / rpb;
  int grid = (int)(want < 148 * 8 ? want : 148 * 8);
  if (grid < 1) grid = 1;
  size_t smem = (size_t)threads * V * sizeof(float);
  if (dtype == 1)
    pool_relu_bwd_bias_grad_kernel<__nv_bfloat16><<<grid, threads, smem, st>>>((const __nv_bfloat16*)gp, (const uint8_t*)idx,
                                                                            (__nv_bfloat16*)gz, gb, N, H, W, cvec);
  else
    pool_relu_bwd_bias_grad_kernel<float><<<grid, threads, smem, st>>>((const float*)gp, (const uint8_t*)idx, (float*)gz, gb, N, H, W, cvec);
  return cudaGetLastError() == cudaSuccess ? 1 : -2;
}
