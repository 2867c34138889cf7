// Head-dimension padding for the tensor-core family.  TMA needs a 16-byte row pitch, i.e. D % 8 == 0 for 16-bit
// operands; the reference covers any D by zero-padding inside its async copies (AttentionKernel+OuterProduct.swift:237-254,
// GEMMHeaders.swift:111-114).  Here an operand whose head dimension is not a multiple of 8 is copied once into a staging
// buffer with pad8(D) columns (zeros in the padding -- exact: every contraction over D only gains zero terms), the tcgen05
// kernels run on the staged operands, and the FP32 outputs are copied back without the padding.  The copies are O(N D)
// against O(N^2 D) of attention work: D = 35, 77, 95, 199 (the reference's own test shapes,
// SquareAttentionTest.swift:6-25) run on the tensor cores instead of the FP32 CUDA-core family.
#include <cuda_runtime.h>
#include <stdint.h>

#include "attention_params.h"
#include "backward_common.cuh"

namespace mfa {
namespace {

// dst[r][c] = c < D ? src[r][c] : 0 for c < Dp; one thread per destination element, rows = batch * seq
template <typename T>
__global__ void __launch_bounds__(256) pad_columns(const T *__restrict__ src, T *__restrict__ dst, uint64_t rows, uint32_t D,
                                                   uint32_t Dp) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * Dp) return;
  const uint64_t r = i / Dp;
  const uint32_t c = static_cast<uint32_t>(i % Dp);
  dst[i] = c < D ? src[r * D + c] : T(0);
}

// dst[r][c] = src[r][c] for c < D (src has Dp columns)
template <typename T>
__global__ void __launch_bounds__(256) unpad_columns(const T *__restrict__ src, T *__restrict__ dst, uint64_t rows, uint32_t D,
                                                     uint32_t Dp) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * D) return;
  const uint64_t r = i / D;
  const uint32_t c = static_cast<uint32_t>(i % D);
  dst[i] = src[r * Dp + c];
}

// dO (BF16) -> FP16, eight elements per thread; exact for 2^-14 <= |x| < 65504 (backward_common.cuh)
__global__ void __launch_bounds__(256) bf16_to_f16(const uint4 *__restrict__ src, uint4 *__restrict__ dst, uint64_t vectors) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < vectors) dst[i] = bwd::bf16x8_to_f16x8(src[i]);
}

}  // namespace

cudaError_t launch_bf16_to_f16(const void *src, void *dst, uint64_t elements, cudaStream_t stream) {
  const uint64_t vectors = elements / 8;  // D % 8 == 0
  bf16_to_f16<<<static_cast<uint32_t>((vectors + 255) / 256), 256, 0, stream>>>(static_cast<const uint4 *>(src),
                                                                                 static_cast<uint4 *>(dst), vectors);
  return cudaGetLastError();
}

cudaError_t launch_pad_columns(const void *src, void *dst, uint64_t rows, uint32_t D, uint32_t Dp, uint32_t element_bytes,
                               cudaStream_t stream) {
  const uint64_t n = rows * Dp;
  const uint32_t blocks = static_cast<uint32_t>((n + 255) / 256);
  if (element_bytes == 2)
    pad_columns<uint16_t><<<blocks, 256, 0, stream>>>(static_cast<const uint16_t *>(src), static_cast<uint16_t *>(dst), rows, D, Dp);
  else
    pad_columns<uint32_t><<<blocks, 256, 0, stream>>>(static_cast<const uint32_t *>(src), static_cast<uint32_t *>(dst), rows, D, Dp);
  return cudaGetLastError();
}

cudaError_t launch_unpad_columns(const void *src, void *dst, uint64_t rows, uint32_t D, uint32_t Dp, cudaStream_t stream) {
  const uint64_t n = rows * D;
  const uint32_t blocks = static_cast<uint32_t>((n + 255) / 256);
  unpad_columns<uint32_t><<<blocks, 256, 0, stream>>>(static_cast<const uint32_t *>(src), static_cast<uint32_t *>(dst), rows, D, Dp);
  return cudaGetLastError();
}

}  // namespace mfa
