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
// All accesses are 16-byte vectors along the contiguous channel dimension.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "cuda/ptx.cuh"

#define BNET_API extern "C" __attribute__((visibility("default")))

namespace bnet {
namespace nn {

template <typename T> struct Vec;
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* f) {
    int4 v = *reinterpret_cast<const int4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; i++) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* f) {
    int4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<int4*>(p) = v;
  }
  // read-once / write-once traffic: do not allocate in L1
  static __device__ __forceinline__ void load_stream(const __nv_bfloat16* p, float* f) {
    int4 v = bnet::ptx::ld_na_v4(reinterpret_cast<const int4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; i++) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
  static __device__ __forceinline__ void store_stream(__nv_bfloat16* p, const float* f) {
    int4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    bnet::ptx::st_na_v4(reinterpret_cast<int4*>(p), v);
  }
};
template <> struct Vec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float* f) {
    float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
  static __device__ __forceinline__ void load_stream(const float* p, float* f) {
    float4 v = bnet::ptx::ld_na_f4(reinterpret_cast<const float4*>(p));
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  static __device__ __forceinline__ void store_stream(float* p, const float* f) {
    bnet::ptx::st_na_f4(reinterpret_cast<float4*>(p), make_float4(f[0], f[1], f[2], f[3]));
  }
};

// ---- y = relu(z + b), in place ------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bias_relu_kernel(T* __restrict__ z, const T* __restrict__ bias, size_t nvec, int cvec) {
  constexpr int V = Vec<T>::N;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float v[V], b[V];
    Vec<T>::load(z + i * V, v);
    Vec<T>::load(bias + (i % cvec) * V, b);
#pragma unroll
    for (int k = 0; k < V; k++) v[k] = fmaxf(v[k] + b[k], 0.f);
    Vec<T>::store(z + i * V, v);
  }
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
// rows = N*H*W, cvec = C / V.  blockDim.x is a multiple of cvec.  U rows per thread are loaded before any
// dependent work so that 2*U 16-byte requests per thread are in flight (HBM-bound streaming kernel).
template <typename T>
__global__ void __launch_bounds__(256) relu_bwd_bias_grad_kernel(const T* __restrict__ gy, const T* __restrict__ y,
                                                                 T* __restrict__ gz, float* __restrict__ gb, size_t rows,
                                                                 int cvec) {
  constexpr int V = Vec<T>::N;
  constexpr int U = 4;
  extern __shared__ float smem[];
  const int rpb = blockDim.x / cvec;
  const int grp = threadIdx.x % cvec, trow = threadIdx.x / cvec;
  const size_t step = (size_t)gridDim.x * rpb;
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; k++) acc[k] = 0.f;
  for (size_t r0 = (size_t)blockIdx.x * rpb + trow; r0 < rows; r0 += U * step) {
    float g[U][V], a[U][V];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t r = r0 + u * step;
      if (r < rows) {
        const size_t off = (r * cvec + grp) * V;
        Vec<T>::load_stream(gy + off, g[u]);
        Vec<T>::load_stream(y + off, a[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t r = r0 + u * step;
      if (r < rows) {
#pragma unroll
        for (int k = 0; k < V; k++) {
          g[u][k] = a[u][k] > 0.f ? g[u][k] : 0.f;
          acc[k] += g[u][k];
        }
        Vec<T>::store_stream(gz + (r * cvec + grp) * V, g[u]);
      }
    }
  }
  reduce_bias_grad<V>(acc, gb, cvec, smem);
}

// ---- p = maxpool2x2(relu(z + b)) with a 1-byte index per element ---------------------------------
// idx: bits 0-1 = position of the max inside the window (dy*2+dx), bit 2 = max > 0 (gradient flows)
template <typename T>
__global__ void __launch_bounds__(256) bias_relu_pool_fwd_kernel(const T* __restrict__ z, const T* __restrict__ bias,
                                                                 T* __restrict__ p, uint8_t* __restrict__ idx, int N, int H,
                                                                 int W, int cvec) {
  constexpr int V = Vec<T>::N;
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)N * Ho * Wo * cvec;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int g = (int)(i % cvec);
    size_t t = i / cvec;
    const int wo = (int)(t % Wo);
    t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const size_t base = (((size_t)n * H + 2 * ho) * W + 2 * wo) * cvec + g;   // in vectors
    float v[4][V], b[V];
    Vec<T>::load(z + base * V, v[0]);
    Vec<T>::load(z + (base + cvec) * V, v[1]);
    Vec<T>::load(z + (base + (size_t)W * cvec) * V, v[2]);
    Vec<T>::load(z + (base + (size_t)W * cvec + cvec) * V, v[3]);
    Vec<T>::load(bias + (size_t)g * V, b);
    float out[V];
    uint8_t code[V];
#pragma unroll
    for (int k = 0; k < V; k++) {
      float m = v[0][k];
      int a = 0;
#pragma unroll
      for (int q = 1; q < 4; q++)
        if (v[q][k] > m) { m = v[q][k]; a = q; }
      m += b[k];
      out[k] = fmaxf(m, 0.f);
      code[k] = (uint8_t)(a | (m > 0.f ? 4 : 0));
    }
    Vec<T>::store(p + i * V, out);
    if constexpr (V == 8) {
      *reinterpret_cast<uint2*>(idx + i * V) = *reinterpret_cast<const uint2*>(code);
    } else {
      *reinterpret_cast<uint32_t*>(idx + i * V) = *reinterpret_cast<const uint32_t*>(code);
    }
  }
}

// ---- gz = scatter(gp through idx) ; gb += sum of routed gradients -----------------------------------
template <typename T>
__global__ void __launch_bounds__(256) pool_relu_bwd_bias_grad_kernel(const T* __restrict__ gp, const uint8_t* __restrict__ idx,
                                                                      T* __restrict__ gz, float* __restrict__ gb, int N, int H,
                                                                      int W, int cvec) {
  constexpr int V = Vec<T>::N;
  constexpr int U = 2;
  extern __shared__ float smem[];
  const int Ho = H / 2, Wo = W / 2;
  const size_t rows = (size_t)N * Ho * Wo;
  const int rpb = blockDim.x / cvec;
  const int g = threadIdx.x % cvec, trow = threadIdx.x / cvec;
  const size_t step = (size_t)gridDim.x * rpb;
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; k++) acc[k] = 0.f;
  for (size_t r0 = (size_t)blockIdx.x * rpb + trow; r0 < rows; r0 += U * step) {
    float gv[U][V];
    uint8_t code[U][V];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t r = r0 + u * step;
      if (r < rows) {
        const size_t i = r * cvec + g;
        Vec<T>::load_stream(gp + i * V, gv[u]);
        if constexpr (V == 8) *reinterpret_cast<uint2*>(code[u]) = *reinterpret_cast<const uint2*>(idx + i * V);
        else *reinterpret_cast<uint32_t*>(code[u]) = *reinterpret_cast<const uint32_t*>(idx + i * V);
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t r = r0 + u * step;
      if (r >= rows) continue;
      size_t t = r;
      const int wo = (int)(t % Wo);
      t /= Wo;
      const int ho = (int)(t % Ho);
      const int n = (int)(t / Ho);
      float o[4][V];
#pragma unroll
      for (int k = 0; k < V; k++) {
        const float gk = (code[u][k] & 4) ? gv[u][k] : 0.f;
        acc[k] += gk;
#pragma unroll
        for (int q = 0; q < 4; q++) o[q][k] = ((code[u][k] & 3) == q) ? gk : 0.f;
      }
      const size_t base = (((size_t)n * H + 2 * ho) * W + 2 * wo) * cvec + g;
      Vec<T>::store_stream(gz + base * V, o[0]);
      Vec<T>::store_stream(gz + (base + cvec) * V, o[1]);
      Vec<T>::store_stream(gz + (base + (size_t)W * cvec) * V, o[2]);
      Vec<T>::store_stream(gz + (base + (size_t)W * cvec + cvec) * V, o[3]);
    }
  }
  reduce_bias_grad<V>(acc, gb, cvec, smem);
}

template <typename T>
static int pick_threads(int cvec) {
  // a multiple of cvec, at most 256
  int t = (256 / cvec) * cvec;
  return t > 0 ? t : cvec;
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
