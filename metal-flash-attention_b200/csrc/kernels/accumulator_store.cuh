// Accumulator epilogue of the backward kernels (tcgen05_backward.cu, tcgen05_backward_generic.cu).
#pragma once
#include <stdint.h>

#include "sm100_ptx.cuh"

namespace mfa {

// TMEM -> global (row-major FP32): TMEM hands every thread one row, and storing rows straight from registers
// touches 32 different cache lines per warp store.  Each warp therefore transposes 32 x 32 chunks through a private
// XOR-swizzled 4 KB scratch tile in shared memory (128-bit accesses, conflict-free both ways) and writes four full
// 128 B lines per store instruction -- the forward kernel's epilogue (tcgen05_forward.cu).  `scratch` overlays the
// staged-operand ring, which is dead once the final commit has arrived.  Columns [col0, col0 + cols) of the
// accumulator at `t_acc` go to rows [warp_row0, warp_row0 + 32) of `out` ([rows_total][D] FP32).
__device__ __forceinline__ void store_accumulator_coalesced(uint32_t t_acc, uint32_t col0, uint32_t cols, float4 *scratch,
                                                            float *out_base, uint32_t warp_row0, uint32_t rows_total,
                                                            uint32_t D, uint32_t lane) {
  const uint32_t sub_row = lane >> 3, quad = lane & 7;  // transposed view: 4 rows x 8 float4 per warp access
  for (uint32_t cc = 0; cc < cols; cc += 32) {
    const uint32_t c = col0 + cc;
    uint32_t o[32];
    ptx::tmem_ld32(t_acc + c, o);
    ptx::tc_wait_ld();
#pragma unroll
    for (uint32_t j = 0; j < 8; ++j)
      scratch[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(o[4 * j]), __uint_as_float(o[4 * j + 1]),
                                                          __uint_as_float(o[4 * j + 2]), __uint_as_float(o[4 * j + 3]));
    __syncwarp();
    // all eight values in distinct registers before the first store (a store holds its source registers until the
    // data has left the SM)
    float4 v[8];
#pragma unroll
    for (uint32_t i = 0; i < 8; ++i) {
      const uint32_t r = 4 * i + sub_row;
      v[i] = scratch[r * 8 + (quad ^ (r & 7))];
    }
    if (c + 4 * quad < D) {  // D % 8 == 0: a float4 is either fully inside or fully outside
#pragma unroll
      for (uint32_t i = 0; i < 8; ++i) {
        const uint32_t r = 4 * i + sub_row;
        if (warp_row0 + r < rows_total)
          *reinterpret_cast<float4 *>(out_base + static_cast<size_t>(r) * D + c + 4 * quad) = v[i];
      }
    }
    __syncwarp();
  }
}

}  // namespace mfa
