// Tensor-core linear layer for sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM -> fused epilogue.
//
//   D[i, j] = sum_k A[i, k] * B[j, k]        A: "lane" operand, 128 rows per tile (one per TMEM lane)
//                                            B: "column" operand, BN rows per tile (the MMA N dimension)
//
// Forward: both operands are K-major bf16 matrices (a torch Linear's activation [M,K] and weight [N,K] as they are), so
// either can take either role: with a large batch x is A and w is B; with a small batch (M <= 64) the roles swap so that
// the 128 TMEM lanes are filled by output features instead of 3/4 padding rows.
// Backward: dX = gY . W reduces over N and dW = gY^T . X reduces over M, so W, gY and X appear with the reduction
// dimension OUTER ("MN-major" operands).  Those are staged as [64 reduction rows x 64 MN elements] TMA boxes (one per 64
// MN elements of the tile, 8 KiB each) and described to the MMA with the MN-major form of the shared-memory descriptor;
// no transposed copy of any tensor is ever made.
//
// Persistent CTAs (one per SM once there are more tiles than SMs), 6 warps, warp-specialised (guide "Anatomy of a
// Blackwell GEMM kernel"); the ring of shared-memory stages runs through tile boundaries:
//   warp 0, one lane   TMA producer: per stage one [128 x 64] lane box and one [BN x 64] column box (K-major operands) or
//                      [64 x 64] boxes per 64 MN elements (MN-major operands), 128-byte swizzle, completion on the
//                      stage's `full` mbarrier; a gathered batch operand is loaded from the owning rank's tensor map
//   warp 1             allocates 2 x BN TMEM columns (two accumulator stages); one lane issues 4 tcgen05.mma (K = 16
//                      each) per stage and tcgen05.commit's the stage's `empty` mbarrier; after a tile's last K block it
//                      commits `acc_full[stage]`, and before reusing a stage it waits for `acc_empty[stage]`
//   warps 2-5          epilogue: tcgen05.ld 32 lanes x 16 columns at a time -> bias / ReLU -> bf16 store, or (reduce
//                      mode) fp32 adds into every rank's output — multimem.red through the NVSwitch multicast mapping
//                      or red.global per peer — or into the owner's output only (reduce-scatter); arrives on
//                      `acc_empty[stage]` as soon as the accumulator is in registers.  Warp w may only touch TMEM
//                      lanes 32*(w%4) .. +31.
// The index logic (tiling, TMA coordinates, epilogue addressing) is csrc/cuda/tc_body.cuh, shared with the CPU
// emulation test; the mbarrier protocol is mirrored by tests/test_tc_pipeline_model.py.
//
// Every mbarrier wait carries a watchdog (kWatchdogNs): a pipeline that stops — a descriptor the hardware rejects, a
// lost TMA completion — sets *err and lets every role fall through to the teardown instead of hanging the GPU.
//
// STATUS: validated on B200 in round 2 (profiles/r2/tc_probe_1gpu.txt: every probe of tools/tc_probe.py, linear forward /
// backward, split-K, GEMM + all-reduce, 3x3 convolution forward + input gradient); the convolution WEIGHT gradient
// (kConv == 2) was written afterwards and has only run on the CPU emulation — ops/tc_conv.py therefore trusts it only after a
// self-check in a child process on the GPU at hand.  Descriptor packing is unit-tested against CuTe (csrc/tests/tc_desc_test.cc).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <mutex>
#include <string>

#include "bnet/bnet_tc.h"
#include "cuda/driver_api.h"
#include "cuda/ptx.cuh"
#include "cuda/tc_body.cuh"

#define BNET_API extern "C" __attribute__((visibility("default")))

namespace {

using namespace bnet;
using namespace bnet::tc;              // tile constants, TcArgs, plan / setup, per-tile index logic (tc_body.cuh)

constexpr int kThreads = 192;
constexpr uint64_t kWatchdogNs = 2000000000ull;

thread_local std::string g_err;

// ------------------------------------------------------------------------------------------------ descriptors
// Instruction descriptor, .kind::f16 (bit positions: cute/arch/mma_sm100_desc.hpp `InstrDescriptor`):
//   [4,6) D format (1 = f32)  [7,10) A format (1 = bf16)  [10,13) B format (1 = bf16)
//   15 / 16 A / B major (0 = K-major)   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t instr_desc_bf16_f32(int m, int n, bool a_mn = false, bool b_mn = false) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(a_mn) << 15) | (uint32_t(b_mn) << 16) | (uint32_t(n >> 3) << 17) |
         (uint32_t(m >> 4) << 24);
}

// Shared-memory matrix descriptor of a K-major tile whose rows are 128 bytes, stored densely and swizzled by the
// hardware in 8-row x 128-byte atoms (what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B):
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4 (unused for swizzled K-major layouts; 1 like CuTe)
//   [32,46) stride byte offset >> 4 = 1024 B between 8-row groups   [46,48) version = 1   [61,64) layout = 2 (SW128)
__host__ __device__ constexpr uint64_t smem_desc_sw128(uint32_t smem_addr) {
  return uint64_t((smem_addr >> 4) & 0x3FFF) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
         (uint64_t(2) << 61);
}

// The same for an MN-major tile staged as [64 reduction rows x 128 bytes] boxes, one per 64 MN elements, 8 KiB apart:
//   leading byte offset = 8192 B between 64-element MN atoms, stride byte offset = 1024 B between 8-row reduction groups
//   (CuTe: make_umma_desc<Major::MN> of tile_to_shape(Layout_MN_SW128_Atom<bf16>, (MN, 64), Step<_2,_1>) -> LBO 512, SBO 64).
// One K = 16 slice is two 8-row groups further: +2048 B.
__host__ __device__ constexpr uint64_t smem_desc_sw128_mn(uint32_t smem_addr) {
  return uint64_t((smem_addr >> 4) & 0x3FFF) | (uint64_t(8192 >> 4) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
         (uint64_t(2) << 61);
}

// ------------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// false = the watchdog fired (and *err now says which role gave up)
__device__ __forceinline__ bool mbar_wait_wd(uint32_t bar, uint32_t parity, int* err, int role) {
  if (mbar_try_wait(bar, parity)) return true;
  const uint64_t t0 = ptx::globaltimer();
  while (!mbar_try_wait(bar, parity)) {
    if (ptx::globaltimer() - t0 > kWatchdogNs) {
      atomicCAS(err, 0, role);
      return false;
    }
  }
  return true;
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" :: "l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]; accumulate == 0 overwrites D
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrives once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// this warp's 32 TMEM lanes x 16 consecutive fp32 columns: v[t] = D[lane, col + t]
__device__ __forceinline__ void tmem_ld_16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void multimem_red_add_f32(float* mc, float v) {
  asm volatile("multimem.red.relaxed.sys.global.add.f32 [%0], %1;" :: "l"(mc), "f"(v) : "memory");
}
__device__ __forceinline__ void multimem_red_add_v4_f32(float* mc, const float4& v) {
  asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ------------------------------------------------------------------------------------------------ kernel
// the epilogue's memory operations (tc_body.cuh::epilogue_chunk)
struct DeviceOut {
  __device__ __forceinline__ void st16(uint16_t* p, uint16_t v) const { *p = v; }
  __device__ __forceinline__ void st16x16(uint16_t* row, const float (&f)[16]) const {
    uint32_t w[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      __nv_bfloat162 p = __floats2bfloat162_rn(f[2 * u], f[2 * u + 1]);
      w[u] = *reinterpret_cast<uint32_t*>(&p);
    }
    reinterpret_cast<uint4*>(row)[0] = make_uint4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<uint4*>(row)[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
  __device__ __forceinline__ void add1(float* p, float v, bool multicast) const {
    if (multicast) multimem_red_add_f32(p, v); else ptx::red_add_f32(p, v);
  }
  __device__ __forceinline__ void add4(float* p, const float* f, bool multicast) const {
    const float4 x = make_float4(f[0], f[1], f[2], f[3]);
    if (multicast) multimem_red_add_v4_f32(p, x); else ptx::red_add_v4_f32(p, x);
  }
};

// the batch operand's tensor maps: one, or one per rank when its row blocks live on different GPUs
struct TcBatchMaps { CUtensorMap m[BNET_TC_MAX_PEERS]; };

template <int BN>
struct Smem {
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

// kSwap:   false -> A = x (lane i = batch row m), B = w (column j = feature n); out[m, n] = out[i * ldo + j]
//          true  -> A = w (lane i = feature n),   B = x (column j = batch row m); out[m, n] = out[j * ldo + i]
// kReduce: fp32 adds into args.outs[] instead of a bf16 store
//
// Persistent over output tiles: CTA c takes tiles c, c + gridDim.x, ... (consecutive CTAs share the column operand's
// block, so it is read from HBM once and from L2 afterwards).  The shared-memory ring runs straight through tile
// boundaries, and the accumulator is double-buffered in TMEM (2 x BN columns): the epilogue of tile t drains one half
// while the MMAs of tile t+1 fill the other.  With gridDim.x == n_tiles every CTA simply does one tile.
// kAMn / kBMn: the lane / column operand is MN-major (reduction dimension outer) instead of K-major
// kConv: 1 = the lane operand is a 3x3 convolution's input patch, loaded as 4-D TMA boxes (tc_body.cuh::ConvGeom);
//        2 = convolution weight gradient: BOTH operands are 4-D boxes of NHWC activations (gy on the lanes through
//            maps_batch.m[0], the shifted input on the columns through map_feat), MN-major, the reduction runs over pixels
template <int BN, int kStages, bool kSwap, bool kReduce, bool kAMn, bool kBMn, int kConv = 0>
__global__ void __launch_bounds__(kThreads, 1)
tc_linear_kernel(const __grid_constant__ TcBatchMaps maps_batch, const __grid_constant__ CUtensorMap map_feat, const TcArgs args) {
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M = 128");
  static_assert((BN & (BN - 1)) == 0 && BN >= 32, "TMEM allocations are powers of two >= 32 columns");
  static_assert(!kBMn || BN % 64 == 0, "an MN-major operand is staged in 64-element atoms");
  static_assert(kConv != 1 || (!kSwap && !kAMn && !kReduce), "convolution: pixels ride the lanes, K-major patches, bf16 store");
  static_assert(kConv != 2 || (!kSwap && kAMn && kBMn && kReduce), "weight gradient: MN-major operands, split over pixel blocks");
  constexpr uint32_t kTmemCols = 2 * BN;          // two accumulator stages
  extern __shared__ uint8_t smem_raw[];
  // swizzle-128B atoms must start on 1024-byte boundaries of the shared window
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* const tiles = smem_raw + (base - raw);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(tiles + kStages * Smem<BN>::kStageBytes);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kStages;
  const uint32_t acc_full0 = empty0 + 8 * kStages, acc_empty0 = acc_full0 + 16;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  uint32_t* const last_slot = tmem_slot + 1;      // split-K fix-up: "this CTA's slice was the last one of the tile"

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_begin = blockIdx.z * args.k_per_split;
  const int kb_end = min(kb_begin + args.k_per_split, args.k_blocks);
  const int nkb = kb_end - kb_begin;       // >= 1 by construction of the grid

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps_batch.m[0]);
    tma_prefetch_desc(&map_feat);
    for (int s = 0; s < kStages; s++) {
      ptx::mbar_init(bars + s, 1);               // full: the producer's arrive.expect_tx (+ TMA bytes)
      ptx::mbar_init(bars + kStages + s, 1);     // empty: one tcgen05.commit
    }
    for (int s = 0; s < 2; s++) {
      ptx::mbar_init(bars + 2 * kStages + s, 1);       // accumulator stage complete: one tcgen05.commit
      ptx::mbar_init(bars + 2 * kStages + 2 + s, 128); // accumulator stage drained: every epilogue thread
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t g = 0;                              // k-blocks issued by this CTA so far (ring position)
      bool alive = true;
      for (int t = blockIdx.x; t < args.n_tiles && alive; t += gridDim.x) {
        // the operand that is the batch rides the lanes (plain) or the columns (swapped); when it is gathered, the tile's
        // rows come from ONE rank's map, at a row offset inside that rank's shard (tc_body.cuh::tile_coord)
        const TileCoord tc = tile_coord<BN, kSwap>(args, t);
        const CUtensorMap* map_a = kSwap ? &map_feat : &maps_batch.m[tc.a_map];
        const CUtensorMap* map_b = kSwap ? &maps_batch.m[tc.b_map] : &map_feat;
        ConvTile ct{};
        if constexpr (kConv == 1) ct = conv_tile<BN>(args, t);
        for (int i = 0; i < nkb; i++, g++) {
          const uint32_t s = g % kStages, round = g / kStages;
          if (round > 0 && !mbar_wait_wd(empty0 + 8 * s, (round - 1) & 1, args.err, 1)) { alive = false; break; }
          const uint32_t a_dst = base + s * Smem<BN>::kStageBytes, bar = full0 + 8 * s;
          mbar_expect_tx(bar, Smem<BN>::kStageBytes);
          if constexpr (kConv == 1) {
            conv_stage_loads<BN, kBMn>(args, ct, kb_begin + i,
                [&](int offset, int c0, int c1, int c2, int c3) { tma_load_4d(a_dst + offset, &maps_batch.m[0], bar, c0, c1, c2, c3); },
                [&](int offset, int c0, int c1) { tma_load_2d(a_dst + kABytes + offset, &map_feat, bar, c0, c1); });
          } else if constexpr (kConv == 2) {
            wgrad_stage_loads<BN>(args, tc, kb_begin + i,
                [&](int offset, int c0, int c1, int c2, int c3) { tma_load_4d(a_dst + offset, &maps_batch.m[0], bar, c0, c1, c2, c3); },
                [&](int offset, int c0, int c1, int c2, int c3) { tma_load_4d(a_dst + kABytes + offset, &map_feat, bar, c0, c1, c2, c3); });
          } else {
            stage_loads<BN, kAMn, kBMn>(tc, (kb_begin + i) * kBK, [&](int operand, int offset, int c0, int c1) {
              tma_load_2d(a_dst + (operand ? kABytes : 0) + offset, operand ? map_b : map_a, bar, c0, c1);
            });
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc_bf16_f32(kBM, BN, kAMn, kBMn);
      // descriptor advance per K = 16 slice, in 16-byte units: 32 B along the swizzled row (K-major), or two 8-row groups
      // = 2048 B (MN-major)
      constexpr uint64_t kStepA = kAMn ? 2048 / 16 : kUmmaK * 2 / 16, kStepB = kBMn ? 2048 / 16 : kUmmaK * 2 / 16;
      uint32_t g = 0, lt = 0;                      // ring position; tiles done by this CTA
      bool alive = true;
      for (int t = blockIdx.x; t < args.n_tiles && alive; t += gridDim.x, lt++) {
        const uint32_t as = lt & 1, use = lt >> 1;
        // the epilogue must have drained this accumulator stage (two tiles ago) before it is overwritten
        if (use > 0 && !mbar_wait_wd(acc_empty0 + 8 * as, (use - 1) & 1, args.err, 2)) break;
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int i = 0; i < nkb; i++, g++) {
          const uint32_t s = g % kStages;
          alive = mbar_wait_wd(full0 + 8 * s, (g / kStages) & 1, args.err, 2);
          if (!alive) break;
          tc_fence_after_sync();
          const uint32_t a_src = base + s * Smem<BN>::kStageBytes;
          const uint64_t da = kAMn ? smem_desc_sw128_mn(a_src) : smem_desc_sw128(a_src);
          const uint64_t db = kBMn ? smem_desc_sw128_mn(a_src + kABytes) : smem_desc_sw128(a_src + kABytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; k++)
            umma_bf16(tmem_d, da + uint64_t(k) * kStepA, db + uint64_t(k) * kStepB, idesc, (i | k) != 0 ? 1u : 0u);
          umma_commit(empty0 + 8 * s);            // the stage may be refilled once these MMAs have read it
        }
        if (alive) umma_commit(acc_full0 + 8 * as);
      }
    }
  } else {
    // ---- epilogue: 4 warps x 32 lanes = the tile's 128 accumulator rows
    const int q = warp & 3;                        // the TMEM lane quarter this warp may read
    const bool fix = kReduce && args.fix_out != nullptr;
    const bool add_bias = args.bias != nullptr && (!kReduce || blockIdx.z == 0) && !fix;
    uint32_t lt = 0;
    for (int t = blockIdx.x; t < args.n_tiles; t += gridDim.x, lt++) {
      const uint32_t as = lt & 1;
      TileCoord tc = tile_coord<BN, kSwap>(args, t);
      int i_glob = tc.a_row0 + q * 32 + lane;         // row of the lane operand this thread owns
      if constexpr (kConv == 1) {
        // the lane is a pixel of the tile's patch: its row in the [N*H*W, Cout] output, or "past the end" outside the image
        const ConvTile ct = conv_tile<BN>(args, t);
        const long long row = conv_out_row(args, ct, q * 32 + lane);
        i_glob = row < 0 ? args.rows_a : (int)row;
        tc.b_row0 = ct.b_row0;
      }
      if (!mbar_wait_wd(acc_full0 + 8 * as, (lt >> 1) & 1, args.err, 3)) break;
      tc_fence_after_sync();
      const uint32_t tmem_d = tmem_base + as * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 16; c++) {
        uint32_t v[16];
        __syncwarp();                                 // tcgen05.ld is .sync.aligned: reconverge after the guarded stores
        tmem_ld_16(tmem_d + (uint32_t(q * 32) << 16) + uint32_t(c * 16), v);
        if (c == BN / 16 - 1) {
          // the whole accumulator stage is in registers: hand it back to the MMA warp before the stores
          tc_fence_before_sync();
          mbar_arrive(acc_empty0 + 8 * as);
        }
        float acc[16];
#pragma unroll
        for (int u = 0; u < 16; u++) acc[u] = __uint_as_float(v[u]);
        epilogue_chunk<kSwap, kReduce>(args, i_glob, tc.b_row0 + c * 16, acc, add_bias, DeviceOut{});
      }
      if constexpr (kReduce) {
        if (fix) {
          // ---- split-K fix-up: the last K slice to arrive finishes the tile (see TcArgs::fix_out)
          __threadfence();                                                   // this thread's adds, device-wide
          asm volatile("bar.sync 1, 128;" ::: "memory");                     // ... and those of all 128 epilogue threads
          if (threadIdx.x == 64) *last_slot = (atomicAdd(args.fix_counters + t, 1) == (int)gridDim.z - 1) ? 1u : 0u;
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (*(volatile uint32_t*)last_slot) {
            __threadfence();
            TcArgs fa = args;
            fa.outs[0] = args.fix_out;
            fa.ldo = args.fix_ldo;
            float* ws = static_cast<float*>(args.outs[0]);
            if (i_glob < args.rows_a) {
#pragma unroll 1
              for (int c = 0; c < BN / 16; c++) {
                const int j0 = tc.b_row0 + c * 16;
                if (j0 >= args.rows_b) break;
                float acc[16];
#pragma unroll
                for (int u = 0; u < 16; u++) {
                  acc[u] = 0.f;
                  if (j0 + u < args.rows_b) {
                    float* p = kSwap ? ws + size_t(j0 + u) * args.ldo + i_glob : ws + size_t(i_glob) * args.ldo + j0 + u;
                    acc[u] = __ldcg(p);
                    __stcg(p, 0.f);                                           // the workspace is clean for the next call
                  }
                }
                epilogue_chunk<kSwap, false>(fa, i_glob, j0, acc, args.bias != nullptr, DeviceOut{});
              }
            }
            if (threadIdx.x == 64) args.fix_counters[t] = 0;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");                     // last_slot is reused by the next tile
        }
      }
    }
  }

  // ---- teardown: every TMEM read and MMA is ordered before the barrier, then the allocating warp frees the columns
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();                       // lane 0 spent the kernel in the MMA loop: tcgen05.dealloc is .sync.aligned
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      (void)cudaGetLastError();
  });
  return fn;
}

// MapDesc (tc_body.cuh) -> CUtensorMap: 2-D bf16, boxes of [box1 x box0], 128-byte swizzle; rows / columns outside the
// matrix read as zero, so ragged extents need no special case in the kernel
bool make_map(CUtensorMap* map, const MapDesc& d) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { g_err = "the CUDA driver does not export cuTensorMapEncodeTiled"; return false; }
  if ((reinterpret_cast<uintptr_t>(d.ptr) & 15) || (size_t(d.pitch_elems) * 2) % 16) {
    g_err = "operands must be 16-byte aligned with a row pitch that is a multiple of 8 elements";
    return false;
  }
  const cuuint64_t dims[2] = {cuuint64_t(d.dim0), cuuint64_t(d.dim1)};
  const cuuint64_t strides[1] = {cuuint64_t(d.pitch_elems) * 2};
  const cuuint32_t box[2] = {cuuint32_t(d.box0), cuuint32_t(d.box1)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d.ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { g_err = "cuTensorMapEncodeTiled failed: " + std::to_string(int(r)); return false; }
  return true;
}

template <int BN, int kStages, bool kSwap, bool kReduce, bool kAMn, bool kBMn, int kConv = 0>
int launch(const TcBatchMaps& mbatch, const CUtensorMap& mfeat, const TcArgs& a, const BnetTcPlan& p, cudaStream_t st) {
  auto kern = tc_linear_kernel<BN, kStages, kSwap, kReduce, kAMn, kBMn, kConv>;
  static std::once_flag once;
  static cudaError_t attr_rc = cudaSuccess;
  const int smem = p.smem_bytes;
  std::call_once(once, [&] { attr_rc = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
  if (attr_rc != cudaSuccess) { g_err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(attr_rc); return -1; }
  kern<<<dim3(p.ctas, 1, p.grid_z), kThreads, smem, st>>>(mbatch, mfeat, a);
  cudaError_t rc = cudaGetLastError();
  if (rc != cudaSuccess) { g_err = std::string("tc_linear launch: ") + cudaGetErrorString(rc); return -1; }
  return 1;
}

// SMs of the current device; 148 (B200) when there is none to ask (host-only planning in the CPU tests)
int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      n = v;
    else { (void)cudaGetLastError(); n = 148; }
  }
  return n;
}

// `batch` / `feat`: the two operands in problem order (out[batch index, feature index]); the plan decides which of them
// rides the TMEM lanes.  bias is indexed by the feature.  All the index logic is tc_body.cuh::setup_problem (pure, unit-
// tested on the CPU); here the map descriptions are encoded and the instantiation is picked.
int run(const Operand& batch, const Operand& feat, int red, const void* bias, void* const* outs, int n_outs, int multicast,
        bool reduce, int ldo, int act, int splits, int* err_dev, void* stream, const void* const* shards = nullptr,
        int n_shards = 0, int scatter_ranks = 0, void* fix_out = nullptr, int* fix_counters = nullptr, int fix_ldo = 0) {
  Problem pr;
  if (const char* e = setup_problem(batch, feat, red, bias, outs, n_outs, multicast, reduce, ldo, act, splits, err_dev, shards,
                                    n_shards, scatter_ranks, sm_count(), &pr)) {
    g_err = e;
    return -1;
  }
  pr.args.fix_out = fix_out;
  pr.args.fix_counters = fix_counters;
  pr.args.fix_ldo = fix_ldo;
  TcBatchMaps mbatch;
  CUtensorMap mfeat;
  for (int r = 0; r < pr.n_batch_maps; r++)
    if (!make_map(&mbatch.m[r], pr.batch_maps[r])) return -1;
  if (!make_map(&mfeat, pr.feat_map)) return -1;
  const BnetTcPlan& p = pr.plan;
  const TcArgs& a = pr.args;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool amn = pr.a_mn, bmn = pr.b_mn, swap = p.swap != 0;
#define BNET_TC_CASE(BN, SWAP, RED, AMN, BMN)                                          \
  if (p.bn == BN && swap == SWAP && reduce == RED && amn == AMN && bmn == BMN)        \
    return launch<BN, stages_for(BN), SWAP, RED, AMN, BMN>(mbatch, mfeat, a, p, st);
  // forward (K-major x K-major), plain and with the cross-rank adds
  BNET_TC_CASE(32, true, false, false, false)   BNET_TC_CASE(32, true, true, false, false)
  BNET_TC_CASE(64, true, false, false, false)   BNET_TC_CASE(64, true, true, false, false)
  BNET_TC_CASE(128, false, false, false, false) BNET_TC_CASE(128, false, true, false, false)
  BNET_TC_CASE(256, false, false, false, false) BNET_TC_CASE(256, false, true, false, false)
  // dX = gY . W: gY K-major; W MN-major — on the columns (large batch) or on the lanes (small batch)
  BNET_TC_CASE(128, false, false, false, true)  BNET_TC_CASE(256, false, false, false, true)
  BNET_TC_CASE(32, true, false, true, false)    BNET_TC_CASE(64, true, false, true, false)
  // dW = gY^T . X: both MN-major
  BNET_TC_CASE(128, false, false, true, true)   BNET_TC_CASE(256, false, false, true, true)
#undef BNET_TC_CASE
  g_err = "no kernel for this plan";
  return -1;
}


// 4-D map over an NHWC activation {C, W, H, N}: boxes of {64 channels, bw, bh, bn}, 128-byte swizzle, zero fill
bool make_map_nhwc(CUtensorMap* map, const MapDesc4& d) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { g_err = "the CUDA driver does not export cuTensorMapEncodeTiled"; return false; }
  if ((reinterpret_cast<uintptr_t>(d.ptr) & 15) || d.C % 8) { g_err = "activations must be 16-byte aligned with C a multiple of 8"; return false; }
  const cuuint64_t dims[4] = {cuuint64_t(d.C), cuuint64_t(d.W), cuuint64_t(d.H), cuuint64_t(d.N)};
  const cuuint64_t strides[3] = {cuuint64_t(d.C) * 2, cuuint64_t(d.W) * d.C * 2, cuuint64_t(d.H) * d.W * d.C * 2};
  const cuuint32_t box[4] = {cuuint32_t(kBK), cuuint32_t(d.bw), cuuint32_t(d.bh), cuuint32_t(d.bn)};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d.ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { g_err = "cuTensorMapEncodeTiled (NHWC) failed: " + std::to_string(int(r)); return false; }
  return true;
}

// the index logic is tc_body.cuh::setup_conv (pure, emulated on the CPU by tc_emu_test); here the maps are encoded and the
// instantiation is picked
int run_conv(const void* x, const void* w, const void* bias, void* out, int N, int H, int W, int Cred, int Cn, int act, int dgrad,
             int* err_dev, void* stream) {
  ConvProblem cp;
  if (const char* e = setup_conv(x, w, bias, out, N, H, W, Cred, Cn, act, dgrad, err_dev, sm_count(), &cp)) {
    g_err = e;
    return -1;
  }
  TcBatchMaps mbatch;
  CUtensorMap mfeat;
  if (!make_map_nhwc(&mbatch.m[0], cp.x_map)) return -1;
  if (!make_map(&mfeat, cp.w_map)) return -1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const BnetTcPlan& p = cp.plan;
#define BNET_CONV_CASE(BN, BMN) \
  if (p.bn == BN && cp.dgrad == BMN) return launch<BN, stages_for(BN), false, false, false, BMN, 1>(mbatch, mfeat, cp.args, p, st);
  BNET_CONV_CASE(64, false) BNET_CONV_CASE(128, false) BNET_CONV_CASE(256, false)
  BNET_CONV_CASE(64, true)  BNET_CONV_CASE(128, true)  BNET_CONV_CASE(256, true)
#undef BNET_CONV_CASE
  g_err = "no convolution kernel for this plan";
  return -1;
}

// filter gradient: index logic tc_body.cuh::setup_conv_wgrad; both tensor maps are 4-D NHWC maps with 64-pixel boxes
int run_conv_wgrad(const void* gy, const void* x, void* dw, float* ws, int* counters, int N, int H, int W, int Cin, int Cout, int splits,
                   int* err_dev, void* stream) {
  WgradProblem wp;
  // BNET_TC_WGRAD_BN=128: every layer on 128-column tiles (ops/tc_conv.py sets it when only that configuration passed the
  // self-check on the GPU at hand)
  // BNET_TC_WGRAD_FIXUP=0: the slices only add into the fp32 workspace, the caller converts and clears it (ops/tc_conv.py)
  static const int force_bn = [] { const char* v = getenv("BNET_TC_WGRAD_BN"); return v ? atoi(v) : 0; }();
  static const int fixup = [] { const char* v = getenv("BNET_TC_WGRAD_FIXUP"); return v ? atoi(v) : 1; }();
  if (const char* e = setup_conv_wgrad(gy, x, ws, dw, counters, N, H, W, Cin, Cout, splits, err_dev, sm_count(), &wp, force_bn, fixup)) {
    g_err = e;
    return -1;
  }
  if (reinterpret_cast<uintptr_t>(dw) & 15) { g_err = "the filter gradient must be 16-byte aligned"; return -1; }
  TcBatchMaps mbatch;
  CUtensorMap mfeat;
  if (!make_map_nhwc(&mbatch.m[0], wp.gy_map)) return -1;
  if (!make_map_nhwc(&mfeat, wp.x_map)) return -1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const BnetTcPlan& p = wp.plan;
  if (p.bn == 128) return launch<128, stages_for(128), false, true, true, true, 2>(mbatch, mfeat, wp.args, p, st);
  if (p.bn == 256) return launch<256, stages_for(256), false, true, true, true, 2>(mbatch, mfeat, wp.args, p, st);
  g_err = "no weight-gradient kernel for this plan";
  return -1;
}

}  // namespace

BNET_API const char* bnet_tc_last_error(void) { return g_err.c_str(); }
BNET_API uint64_t bnet_tc_smem_desc(uint32_t smem_addr) { return smem_desc_sw128(smem_addr); }
BNET_API uint64_t bnet_tc_smem_desc_mn(uint32_t smem_addr) { return smem_desc_sw128_mn(smem_addr); }
BNET_API uint32_t bnet_tc_instr_desc(int m, int n) { return instr_desc_bf16_f32(m, n); }
BNET_API uint32_t bnet_tc_instr_desc2(int m, int n, int a_mn, int b_mn) { return instr_desc_bf16_f32(m, n, a_mn != 0, b_mn != 0); }

BNET_API int bnet_tc_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return (major == 10 && encode_fn() != nullptr) ? 1 : 0;
}

BNET_API int bnet_tc_plan(int M, int N, int K, int reduce, int splits, BnetTcPlan* p) {
  if (!p || M < 1 || N < 1 || K < 1) { g_err = "bad problem size"; return -1; }
  if (K % 8) { g_err = "K must be a multiple of 8 (16-byte TMA row pitch)"; return -1; }
  if (plan_gemm(M, N, K, reduce, splits, sm_count(), p) != 0) { g_err = "problem too large for one launch"; return -1; }
  return 0;
}

BNET_API int bnet_tc_linear(const void* x, const void* w, const void* bias, void* out, int M, int N, int K, int ldx, int ldw,
                            int ldo, int act, int* err_dev, void* stream) {
  void* outs[1] = {out};
  return run(Operand{x, M, ldx, 0}, Operand{w, N, ldw, 0}, K, bias, outs, 1, 0, false, ldo, act, 1, err_dev, stream);
}

BNET_API int bnet_tc_linear_reduce(const void* x, const void* w, const void* bias, void* const* outs, int n_outs, int multicast,
                                   int M, int N, int K, int ldx, int ldw, int ldo, int splits, int* err_dev, void* stream) {
  return run(Operand{x, M, ldx, 0}, Operand{w, N, ldw, 0}, K, bias, outs, n_outs, multicast, true, ldo, BNET_TC_ACT_NONE, splits,
             err_dev, stream);
}

// Split-K with the finish inside the kernel: `ws` is an fp32 [M, N] workspace and `counters` one int per output tile
// (bnet_tc_plan: grid_x * grid_y), BOTH all zero on entry — the kernel leaves them all zero again, so they are allocated
// and cleared once.  out = act(x . w^T + bias) in bf16, one launch.
BNET_API int bnet_tc_linear_splitk(const void* x, const void* w, const void* bias, void* out, float* ws, int* counters, int M, int N,
                                   int K, int ldx, int ldw, int ldo, int act, int splits, int* err_dev, void* stream) {
  void* outs[1] = {ws};
  if (!ws || !counters) { g_err = "split-K needs a workspace and tile counters"; return -1; }
  return run(Operand{x, M, ldx, 0}, Operand{w, N, ldw, 0}, K, bias, outs, 1, 0, true, N, act, splits, err_dev, stream, nullptr, 0, 0,
             out, counters, ldo);
}

// dx[M, K] = gy[M, N] . w[N, K]: the reduction runs over N, so w is read with its rows as the reduction (MN-major)
BNET_API int bnet_tc_linear_dgrad(const void* gy, const void* w, void* dx, int M, int N, int K, int ldgy, int ldw, int lddx,
                                  int* err_dev, void* stream) {
  void* outs[1] = {dx};
  return run(Operand{gy, M, ldgy, 0}, Operand{w, K, ldw, 1}, N, nullptr, outs, 1, 0, false, lddx, BNET_TC_ACT_NONE, 1, err_dev,
             stream);
}

// dw[N, K] = gy[M, N]^T . x[M, K]: the reduction runs over the batch; both operands are MN-major.  dw's two extents are
// feature counts, so the plan never swaps when N > 64 (and a layer with <= 64 outputs simply takes the swapped tiles).
BNET_API int bnet_tc_linear_wgrad(const void* gy, const void* x, void* dw, int M, int N, int K, int ldgy, int ldx, int lddw,
                                  int* err_dev, void* stream) {
  void* outs[1] = {dw};
  if (N <= 64) { g_err = "wgrad needs more than 64 output features (use cuBLAS for tiny layers)"; return -1; }
  return run(Operand{gy, N, ldgy, 1}, Operand{x, K, ldx, 1}, M, nullptr, outs, 1, 0, false, lddw, BNET_TC_ACT_NONE, 1, err_dev,
             stream);
}

// out[n_shards * rows_per_shard, N] = act([x_0; x_1; ...] . w^T + bias): the all-gather of a sequence/batch-sharded
// activation fused into the GEMM's operand loads — shard r is read straight from x_shards[r] (rank r's symmetric heap,
// mapped over NVLink) by TMA, tile by tile, while the tensor cores work on the previous tile.
BNET_API int bnet_tc_allgather_linear(const void* const* x_shards, int n_shards, int rows_per_shard, const void* w, const void* bias,
                                      void* out, int N, int K, int ldx, int ldw, int ldo, int act, int* err_dev, void* stream) {
  void* outs[1] = {out};
  if (n_shards < 1 || rows_per_shard < 1) { g_err = "bad shard arguments"; return -1; }
  return run(Operand{x_shards[0], n_shards * rows_per_shard, ldx, 0}, Operand{w, N, ldw, 0}, K, bias, outs, 1, 0, false, ldo, act, 1,
             err_dev, stream, x_shards, n_shards, 0);
}

// outs[r][M / n_ranks, N] += rows [r * M / n_ranks, (r + 1) * M / n_ranks) of x . w^T: GEMM + reduce-scatter in one kernel
// (every rank calls it with its K-shard; each output tile travels once, to its owner).
BNET_API int bnet_tc_linear_reduce_scatter(const void* x, const void* w, const void* bias, void* const* outs, int n_ranks, int M,
                                           int N, int K, int ldx, int ldw, int ldo, int splits, int* err_dev, void* stream) {
  return run(Operand{x, M, ldx, 0}, Operand{w, N, ldw, 0}, K, bias, outs, n_ranks, 0, true, ldo, BNET_TC_ACT_NONE, splits, err_dev,
             stream, nullptr, 0, n_ranks);
}

// 3x3 / stride 1 / pad 1 convolution on the same kernel (implicit GEMM: 4-D TMA boxes of the NHWC activation, the padding
// is the TMA unit's zero fill), bias and ReLU in the epilogue.  bf16 NHWC in and out, filters as torch stores a
// channels_last Conv2d weight ([Cout][3][3][Cin]).  Cin must be a multiple of 64, Cout of 8.
BNET_API int bnet_tc_conv3x3(const void* x, const void* w, const void* bias, void* out, int N, int H, int W, int Cin, int Cout,
                             int act, int* err_dev, void* stream) {
  return run_conv(x, w, bias, out, N, H, W, Cin, Cout, act, 0, err_dev, stream);
}
// dx[N,H,W,Cin] = conv3x3 of gy[N,H,W,Cout] with the same filter w[Cout][3][3][Cin], taps flipped, filter read MN-major
// (no transposed / rotated copy of the filter is made).  Cout and Cin must be multiples of 64.
BNET_API int bnet_tc_conv3x3_dgrad(const void* gy, const void* w, void* dx, int N, int H, int W, int Cin, int Cout, int* err_dev,
                                   void* stream) {
  return run_conv(gy, w, nullptr, dx, N, H, W, Cout, Cin, BNET_TC_ACT_NONE, 1, err_dev, stream);
}

// dw[Cout][3][3][Cin] = filter gradient of the same convolution: gy [N,H,W,Cout] and x [N,H,W,Cin] are read as 4-D TMA boxes of
// 64 pixels (the reduction), both MN-major; the pixel blocks are split over grid.z and the slice that arrives last at a tile
// converts the fp32 sums to bf16 (split-K fix-up).  `ws` (fp32 [Cout][9 Cin]) and `counters` (bnet_tc_conv3x3_wgrad_tiles ints)
// must be all zero on entry and are all zero again on exit.  splits <= 0: automatic.  Cin and Cout multiples of 64.
BNET_API int bnet_tc_conv3x3_wgrad(const void* gy, const void* x, void* dw, float* ws, int* counters, int N, int H, int W, int Cin,
                                   int Cout, int splits, int* err_dev, void* stream) {
  return run_conv_wgrad(gy, x, dw, ws, counters, N, H, W, Cin, Cout, splits, err_dev, stream);
}
BNET_API int bnet_tc_conv3x3_wgrad_tiles(int Cin, int Cout) { return wgrad_max_tiles(Cin, Cout); }
// the tiling of a filter gradient (pure host function, like bnet_tc_plan): grid_x x grid_y tiles of 128 x bn, grid_z slices of
// the k_blocks 64-pixel blocks, ctas per slice
BNET_API int bnet_tc_conv3x3_wgrad_plan(int N, int H, int W, int Cin, int Cout, int splits, BnetTcPlan* plan) {
  if (!plan) { g_err = "null plan"; return -1; }
  WgradProblem wp;
  float ws_dummy = 0.f;
  int counters_dummy = 0, err_dummy = 0;
  if (const char* e = setup_conv_wgrad(&ws_dummy, &ws_dummy, &ws_dummy, &ws_dummy, &counters_dummy, N, H, W, Cin, Cout, splits,
                                       &err_dummy, sm_count(), &wp)) {
    g_err = e;
    return -1;
  }
  *plan = wp.plan;
  return 0;
}
