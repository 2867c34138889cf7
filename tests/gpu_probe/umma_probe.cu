// Stand-alone hardware probe for the tcgen05 building blocks the attention kernels rely on.
// It checks, against a CPU reference, on one CTA:
//   mode 0  SS MMA, A K-major x B K-major     (S = Q K^T)
//   mode 1  SS MMA, A K-major x B MN-major    (O = P V with P staged in shared memory)
//   mode 2  TS MMA, A in TMEM   x B MN-major  (O = P V with P written to TMEM by tcgen05.st)
// Descriptor fields (LBO, SBO, per-k-step advance) are kernel arguments so the host can sweep
// alternatives in one GPU session when the expected encoding turns out wrong.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I../../metal-flash-attention_b200/csrc/kernels \
//        umma_probe.cu ../../metal-flash-attention_b200/csrc/kernels/tma_host.cpp -o _build/umma_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "sm100_ptx.cuh"
#include "attention_params.h"
#include "tma_host.h"

using namespace mfa;
using namespace mfa::ptx;

struct ProbeArgs {
  int mode;
  uint32_t a_lbo, a_sbo, a_kstep;  // bytes
  uint32_t b_lbo, b_sbo, b_kstep;  // bytes
  uint32_t b_mn_major;
  uint32_t N;                      // MMA N (and columns of D)
  uint32_t ksteps;                 // number of K=16 MMAs
  uint32_t b_format;               // 1 = BF16 (default), 0 = FP16 (B buffer holds FP16 data): mixed A/B formats
};

// A: [128][128] bf16 (K-major rows), B: [128][128] bf16. Both arrive as two [128][64] SW128 sub-tiles.
__global__ void __launch_bounds__(128, 1)
    probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const __nv_bfloat16 *__restrict__ Aglobal, float *__restrict__ Dout, ProbeArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sA = smem, *sB = smem + 32768;
  uint64_t *bar_load = reinterpret_cast<uint64_t *>(smem + 65536);
  uint64_t *bar_mma = bar_load + 1;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bar_load + 2);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_load, 65536);
    for (int ds = 0; ds < 2; ++ds) {
      tma_load_3d(sA + ds * 16384, &mapA, bar_load, ds * 64, 0, 0);
      tma_load_3d(sB + ds * 16384, &mapB, bar_load, ds * 64, 0, 0);
    }
  }
  mbar_wait(bar_load, 0);

  // mode 2: stage A (as "P") into TMEM columns [256, 320): lane = row, column c holds k = 2c, 2c+1
  const uint32_t row = warp * 32 + lane;
  const uint32_t lane_addr = (warp * 32) << 16;
  if (args.mode == 2) {
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t packed[16];
      for (int i = 0; i < 16; ++i) {
        float lo = __bfloat162float(Aglobal[row * 128 + 2 * (c0 + i)]);
        float hi = __bfloat162float(Aglobal[row * 128 + 2 * (c0 + i) + 1]);
        packed[i] = pack_bf16x2(lo, hi);
      }
      tmem_st16(tmem_base + lane_addr + 256 + c0, packed);
    }
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_f16_mixed(128, args.N, 1, args.b_format, 0, args.b_mn_major);
    for (uint32_t k = 0; k < args.ksteps; ++k) {
      // K-major operands: 4 k-steps of 32 B inside a 64-element sub-tile, then the next sub-tile (16 KiB)
      const uint32_t a_off = (k >> 2) * 16384 + (k & 3) * args.a_kstep;
      const uint32_t b_off = args.b_mn_major ? k * args.b_kstep : (k >> 2) * 16384 + (k & 3) * args.b_kstep;
      const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB) + b_off, args.b_lbo, args.b_sbo);
      if (args.mode == 2) {
        umma_ts(tmem_base, tmem_base + 256 + k * 8, bdesc, idesc, k > 0);
      } else {
        const uint64_t adesc = make_smem_desc_sw128(smem_u32(sA) + a_off, args.a_lbo, args.a_sbo);
        umma_ss(tmem_base, adesc, bdesc, idesc, k > 0);
      }
    }
    umma_commit(bar_mma);
  }
  mbar_wait(bar_mma, 0);
  tc_fence_after();

  for (uint32_t c = 0; c < args.N; c += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + lane_addr + c, v);
    tc_wait_ld();
    for (int i = 0; i < 32; ++i) Dout[row * args.N + c + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  const int M = 128, K = 128, N = 128;
  std::vector<float> A(M * K), B(N * K);
  std::vector<__nv_bfloat16> Ab(M * K), Bb(N * K);
  srand(1);
  for (int i = 0; i < M * K; ++i) { A[i] = bf16_round((rand() % 2001 - 1000) / 1000.0f); Ab[i] = __float2bfloat16(A[i]); }
  for (int i = 0; i < N * K; ++i) { B[i] = bf16_round((rand() % 2001 - 1000) / 1000.0f); Bb[i] = __float2bfloat16(B[i]); }
  // reference 0: D0[m][n] = sum_k A[m][k] * B[n][k]      (B rows are N, K-major)
  // reference 1: D1[m][n] = sum_k A[m][k] * B[k][n]      (B rows are K, MN-major; needs N == K == 128)
  std::vector<float> D0(M * N), D1(M * N);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s0 = 0, s1 = 0;
      for (int k = 0; k < K; ++k) { s0 += (double)A[m * K + k] * B[n * K + k]; s1 += (double)A[m * K + k] * B[k * N + n]; }
      D0[m * N + n] = (float)s0; D1[m * N + n] = (float)s1;
    }

  __nv_bfloat16 *dA, *dB; float *dD;
  cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, Ab.data(), M * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bb.data(), N * K * 2, cudaMemcpyHostToDevice);
  CUtensorMap mapA, mapB;
  // FP16 copy of B for the mixed-format checks
  std::vector<__half> Bh(N * K);
  for (int i = 0; i < N * K; ++i) Bh[i] = __float2half(B[i]);
  __half *dBh; cudaMalloc(&dBh, N * K * 2);
  cudaMemcpy(dBh, Bh.data(), N * K * 2, cudaMemcpyHostToDevice);
  CUtensorMap mapBh;
  if (make_tensor_map_16bit(&mapA, dA, 128, 128, 1, 128) != cudaSuccess ||
      make_tensor_map_16bit(&mapBh, dBh, 128, 128, 1, 128) != cudaSuccess ||
      make_tensor_map_16bit(&mapB, dB, 128, 128, 1, 128) != cudaSuccess) {
    printf("PROBE tensor map failed: %s\n", last_launch_detail());
    return 2;
  }
  const int smem_bytes = 65536 + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);

  struct Candidate { const char *name; ProbeArgs a; };
  std::vector<Candidate> cands = {
      {"mode0 SS K-major x K-major (expected encoding)", {0, 16, 1024, 32, 16, 1024, 32, 0, 128, 8, 1}},
      {"mode1 SS K-major x MN-major lbo=16384 sbo=1024 kstep=2048 (expected)", {1, 16, 1024, 32, 16384, 1024, 2048, 1, 128, 8, 1}},
      {"mode2 TS TMEM-A x MN-major (expected)", {2, 16, 1024, 32, 16384, 1024, 2048, 1, 128, 8, 1}},
      {"mode1 N=64 (single column block)", {1, 16, 1024, 32, 16384, 1024, 2048, 1, 64, 8, 1}},
      {"mode2 N=64 (single column block)", {2, 16, 1024, 32, 16384, 1024, 2048, 1, 64, 8, 1}},
      // Measured on B200 (round 1): a_format != b_format under kind::f16 ("MIXED A=bf16 B=fp16") raises
      // "illegal instruction" -- FP16 and BF16 operands cannot be mixed in one tcgen05.mma.  Not run by default
      // because the fault kills the context:
      // {"mode0 MIXED A=bf16 B=fp16 K-major", {0, 16, 1024, 32, 16, 1024, 32, 0, 128, 8, 0}},
  };
  int failures = 0;
  std::vector<float> D(M * N);
  for (auto &c : cands) {
    cudaMemset(dD, 0xFF, M * N * 4);
    probe_kernel<<<1, 128, smem_bytes>>>(mapA, c.a.b_format == 0 ? mapBh : mapB, dA, dD, c.a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("PROBE %-70s CUDA ERROR %s\n", c.name, cudaGetErrorString(e)); return 3; }
    cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    const std::vector<float> &ref = c.a.mode == 0 ? D0 : D1;
    double maxerr = 0; int nan = 0;
    for (int m = 0; m < M; ++m)
      for (uint32_t n = 0; n < c.a.N; ++n) {
        float got = D[m * c.a.N + n], want = ref[m * N + n];
        if (isnan(got)) nan++; else maxerr = fmax(maxerr, fabs(got - want));
      }
    bool ok = nan == 0 && maxerr < 1e-2;
    const bool expected = strstr(c.name, "MIXED") == nullptr;
    if (expected && !ok) failures++;
    printf("PROBE %-70s max_err=%.4g nan=%d %s\n", c.name, maxerr, nan, ok ? "OK" : (expected ? "FAIL" : "(informational)"));
  }
  printf("PROBE SUMMARY failures=%d\n", failures);
  return failures ? 1 : 0;
}
