// Operation codes of the transport executor (csrc/cuda/exec_body.cuh implements them, csrc/cuda/nvl_exec.cu runs them):
// what happens to a message on its way into the receiver's buffer.
#pragma once
#include <stdint.h>

namespace bnet {
namespace cuda {

enum ExecOp : uint32_t {
  OP_COPY = 0,
  OP_RED_ADD_F32 = 1,      // dst(f32) += src(f32)           (K4: accumulate while moving)
  OP_RED_ADD_BF16 = 2,     // dst(bf16) += src(bf16)
  OP_CAST_BF16_TO_F32 = 3, // dst(f32) = src(bf16)           (K5)
  OP_CAST_F32_TO_BF16 = 4, // dst(bf16) = src(f32)
  OP_FLUSH = 5,            // K7: fence only
  OP_ACC_BF16_TO_F32 = 6,  // dst(f32) += src(bf16)          (K4+K5 fused)
  OP_CAST_BF16_TO_E4M3 = 7,  // dst(fp8 e4m3) = sat(src(bf16) * scale)   (gradient compression, K5)
  OP_ACC_E4M3_TO_F32 = 8,    // dst(f32) += src(fp8 e4m3) * scale        (decompress + accumulate)
  OP_CAST_F32_TO_E4M3 = 9,   // dst(fp8 e4m3) = sat(src(f32) * scale)
  OP_CAST_BF16_TO_E5M2 = 10, // the same three with the wide-range e5m2 format
  OP_ACC_E5M2_TO_F32 = 11,
  OP_CAST_F32_TO_E5M2 = 12,
  OP_CAST_E4M3_TO_F32 = 13,  // dst(f32) = src(fp8 e4m3) * scale         (decompress, overwriting: the all-gather half of a
  OP_CAST_E5M2_TO_F32 = 14,  //                                            compressed all-reduce)
  OP_COUNT = 15,
};

}  // namespace cuda
}  // namespace bnet
