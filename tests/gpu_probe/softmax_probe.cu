// Micro-benchmark of the forward kernel's softmax step in isolation (no TMA, no MMA, no mbarriers):
// TMEM S -> registers -> exp2 / row sum / 16-bit pack -> TMEM P, repeated, timed with clock64().
// Answers one design question: how long does one 128 x 128 tile-block of softmax take when
//   mode 0: 4 warps, one thread per row (128 columns per thread, two 64-column halves)       [production, one tile]
//   mode 1: 8 warps, two tiles side by side, one thread per row                               [production, both tiles]
//   mode 2: 8 warps on ONE tile, two threads per row (64 columns each) + bar.red.or consensus on the overflow check
//   mode 3: as 2 without the consensus barrier
// Build: make -C tests/gpu_probe   Run (GPU box): tests/gpu_probe/_build/softmax_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sm100_ptx.cuh"

using namespace mfa::ptx;

template <int POLY, bool BF16>
__device__ __forceinline__ float exp_pack_64(const float *s, float scale_log2, float m, uint32_t *packed) {
  float2 sum2 = make_float2(0.f, 0.f);
  const float2 scale2 = make_float2(scale_log2, scale_log2), negm2 = make_float2(-m, -m);
#pragma unroll
  for (uint32_t i = 0; i < 32; ++i) {
    const float2 x = ffma2(make_float2(s[2 * i], s[2 * i + 1]), scale2, negm2);
    float2 pr;
    if (POLY > 0 && (i & 3) < POLY) {
      pr = exp2_poly2(x);
    } else {
      pr.x = ex2_approx(x.x);
      pr.y = ex2_approx(x.y);
    }
    sum2 = fadd2(sum2, pr);
    packed[i] = BF16 ? pack_bf16x2(pr.x, pr.y) : pack_f16x2(pr.x, pr.y);
  }
  return sum2.x + sum2.y;
}

__device__ __forceinline__ bool pair_any(bool flag, uint32_t barrier_id) {
  uint32_t out;
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.u32 p, %1, 0;\n"
      "bar.red.or.pred q, %2, 64, p;\n"
      "selp.u32 %0, 1, 0, q;\n"
      "}\n"
      : "=r"(out)
      : "r"(static_cast<uint32_t>(flag)), "r"(barrier_id)
      : "memory");
  return out != 0;
}

template <int MODE, int POLY>
__global__ void __launch_bounds__(MODE == 0 ? 128 : 256, 1)
    softmax_probe(float *sink, long long *cycles, int iters, float scale_log2, float m) {
  extern __shared__ uint8_t smem[];
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(smem);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t lane_addr = ((warp & 3) * 32) << 16;

  // S for both tiles: columns 0-255, values in [-3, 3) from a hash of (row, column); P goes to columns 256-383
  {
    const uint32_t row = (warp & 3) * 32 + lane;
    for (uint32_t c0 = (warp >> 2) * 128; c0 < 256; c0 += (blockDim.x >= 256 ? 256 : 128)) {
      for (uint32_t c = 0; c < 128; c += 32) {
        uint32_t v[32];
#pragma unroll
        for (uint32_t i = 0; i < 32; ++i) {
          uint32_t h = (row * 131u + (c0 + c + i) * 2654435761u) ^ (blockIdx.x * 97u);
          h ^= h >> 13;
          h *= 0x5bd1e995u;
          h ^= h >> 15;
          v[i] = __float_as_uint((static_cast<float>(h & 0xffff) / 65536.0f - 0.5f) * 6.0f);
        }
        tmem_st32(tmem_base + lane_addr + c0 + c, v);
      }
    }
    if (MODE == 0) {  // 4 warps: also fill tile 1's columns so that every mode starts from the same state
      for (uint32_t c = 128; c < 256; c += 32) {
        uint32_t v[32];
#pragma unroll
        for (uint32_t i = 0; i < 32; ++i) v[i] = 0;
        tmem_st32(tmem_base + lane_addr + c, v);
      }
    }
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  float l = 0.f;
  const long long t0 = clock64();
  if (MODE == 0 || MODE == 1) {
    const uint32_t t = warp >> 2;
    const uint32_t tS = tmem_base + lane_addr + t * 128;
    const uint32_t tP = tmem_base + lane_addr + 256 + t * 64;
    for (int it = 0; it < iters; ++it) {
      float s[128];
#pragma unroll
      for (uint32_t c = 0; c < 128; c += 32) tmem_ld32(tS + c, *reinterpret_cast<uint32_t(*)[32]>(&s[c]));
      tc_wait_ld();
#pragma unroll
      for (uint32_t half = 0; half < 2; ++half) {
        uint32_t packed[32];
        float half_sum = exp_pack_64<POLY, true>(&s[half * 64], scale_log2, m, packed);
        if (__any_sync(0xffffffffu, !(half_sum <= 256.0f))) {  // never taken with this data; keeps the code shape
          m += 1.0f;
          half_sum = exp_pack_64<0, true>(&s[half * 64], scale_log2, m, packed);
        }
        l += half_sum;
        tmem_st32(tP + half * 32, packed);
        tc_wait_st();
        tc_fence_before();
      }
    }
  } else {
    const uint32_t h = warp >> 2;  // column half of the row this thread owns
    const uint32_t tS = tmem_base + lane_addr + h * 64;
    const uint32_t tP = tmem_base + lane_addr + 256 + h * 32;
    for (int it = 0; it < iters; ++it) {
      float s[64];
#pragma unroll
      for (uint32_t c = 0; c < 64; c += 32) tmem_ld32(tS + c, *reinterpret_cast<uint32_t(*)[32]>(&s[c]));
      tc_wait_ld();
      uint32_t packed[32];
      float half_sum = exp_pack_64<POLY, true>(s, scale_log2, m, packed);
      bool over = __any_sync(0xffffffffu, !(half_sum <= 256.0f));
      if (MODE == 2) over = pair_any(over, 1 + (warp & 3));
      if (over) {
        m += 1.0f;
        half_sum = exp_pack_64<0, true>(s, scale_log2, m, packed);
      }
      l += half_sum;
      tmem_st32(tP, packed);
      tc_wait_st();
      tc_fence_before();
    }
  }
  const long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = l + m;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int MODE, int POLY>
static void run(const char *name, int tiles_per_step) {
  const int ctas = 148, iters = 2000, threads = MODE == 0 ? 128 : 256;
  const size_t smem = 200 * 1024;
  float *sink;
  long long *cycles;
  cudaMalloc(&sink, ctas * threads * sizeof(float));
  cudaMalloc(&cycles, ctas * sizeof(long long));
  auto kernel = softmax_probe<MODE, POLY>;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  for (int rep = 0; rep < 2; ++rep) kernel<<<ctas, threads, smem>>>(sink, cycles, iters, 0.1275f, 0.5f);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s: CUDA error %s\n", name, cudaGetErrorString(e));
    exit(1);
  }
  std::vector<long long> h(ctas);
  cudaMemcpy(h.data(), cycles, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
  double sum = 0;
  for (long long c : h) sum += static_cast<double>(c);
  const double per_step = sum / ctas / iters;
  printf("%-58s poly %d/4: %7.0f cycles/step = %7.0f cycles per 128x128 tile-block\n", name, POLY, per_step,
         per_step / tiles_per_step);
  cudaFree(sink);
  cudaFree(cycles);
}

int main() {
  run<0, 0>("4 warps, 1 thread/row, one tile (latency of a tile step)", 1);
  run<0, 1>("4 warps, 1 thread/row, one tile (latency of a tile step)", 1);
  run<0, 2>("4 warps, 1 thread/row, one tile (latency of a tile step)", 1);
  run<1, 0>("8 warps, 1 thread/row, two tiles at once (throughput)", 2);
  run<1, 1>("8 warps, 1 thread/row, two tiles at once (throughput)", 2);
  run<1, 2>("8 warps, 1 thread/row, two tiles at once (throughput)", 2);
  run<2, 0>("8 warps, 2 threads/row + bar.red consensus, one tile", 1);
  run<2, 1>("8 warps, 2 threads/row + bar.red consensus, one tile", 1);
  run<2, 2>("8 warps, 2 threads/row + bar.red consensus, one tile", 1);
  run<3, 0>("8 warps, 2 threads/row, no consensus, one tile", 1);
  run<3, 1>("8 warps, 2 threads/row, no consensus, one tile", 1);
  return 0;
}
