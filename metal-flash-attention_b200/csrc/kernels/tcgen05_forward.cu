// FlashAttention forward for sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM.
//
// Replaces the reference's generated forward kernel (loopForward, AttentionKernel+Source.swift:158-200;
// outer product S = Q K^T, +OuterProduct.swift:18-487; online softmax, +Softmax.swift:228-324,334-505;
// accumulate O += P V, +Accumulate.swift:24-582) for 16-bit row-major operands.
//
// One CTA owns one 128-row tcgen05 M-tile of Q and walks the keys in blocks of 128.  Warp roles (384 threads):
//   warps 0-3   softmax, columns  0-63  of every S block (thread = one query row = one TMEM lane)
//   warps 4-7   softmax, columns 64-127 of every S block (same rows; warp w and w+4 form a "row pair")
//   warp  8     MMA issuer (one elected lane issues every tcgen05.mma / commit); owns the TMEM allocation
//   warp  9     TMA producer for Q (once) and the K stages;  warp 10  TMA producer for the V stages
//   warp  11    idle (the whole producer warpgroup donates registers via setmaxnreg)
// On-chip residency (the reference's "cache Q, O" rows, AttentionDescriptor+Parameters.swift:109-120,
// re-expressed for B200): Q stays in SMEM for the whole traversal and the O accumulator stays in TMEM.  S is
// TRIPLE-buffered in TMEM and overwritten in place by P (16-bit), which feeds the second MMA straight from
// TMEM:  columns [0,128) [128,256) [256,384) S/P buffers, [384,384+D) O.  S(i+1) is therefore complete before
// the softmax warps start block i, which lets every thread software-pipeline: while the exp2 stream of block i
// occupies the MUFU pipe, the same thread loads S(i+1), reduces its row max on the ALU pipe and swaps the
// half-row maxima with its pair warp through shared memory.  Tensor pipe (8 + 8 MMAs per block) and MUFU pipe
// (128 x 128 exp2 per block) then both run continuously.
// Softmax bookkeeping follows Appendix A of SURVEY.md (log2 domain, L = m + log2 l) with one B200-specific
// change: the running max is only refreshed when it grows by more than 2^8 ("lazy rescale"), so the
// O *= correction pass over TMEM is rare; results are mathematically identical.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include <mutex>

#include "attention_params.h"
#include "sm100_ptx.cuh"
#include "tma_host.h"

namespace mfa {
namespace fwd {

using namespace ptx;

constexpr uint32_t kTileM = 128;   // query rows per CTA (one tcgen05 M-tile)
constexpr uint32_t kBlockN = 128;  // keys per traversal block
constexpr uint32_t kHalfN = 64;    // S columns per softmax warpgroup
constexpr uint32_t kHalves = 2;
constexpr uint32_t kSBuffers = 3;  // S/P buffers in TMEM
constexpr uint32_t kSubTileBytes = 128 * 128;  // [128 rows][64 x 16-bit]: one 128B-swizzled TMA box
constexpr uint32_t kThreads = 384;
constexpr uint32_t kSoftmaxThreads = 256;
// setmaxnreg budget: the CTA is launched with floor(65536 / 384 / 8) * 8 = 168 registers per thread; the two
// softmax warpgroups grow to kSoftmaxRegs after the producer warpgroup has shrunk to kOtherRegs.  The sum
// must not exceed the launch allocation or the second setmaxnreg.inc never returns.
constexpr uint32_t kLaunchRegs = 168, kSoftmaxRegs = 208, kOtherRegs = 88;
static_assert(kSoftmaxRegs * 256 + kOtherRegs * 128 <= kLaunchRegs * kThreads, "setmaxnreg over-subscribed");
constexpr float kRescaleThreshold = 8.0f;  // log2 units

template <uint32_t DPAD>
struct Config {
  static constexpr uint32_t kSubTiles = DPAD / 64;                   // 64-element sub-tiles along D
  static constexpr uint32_t kTileBytes = kSubTiles * kSubTileBytes;  // one 128 x DPAD operand tile
  static constexpr uint32_t kStagesK = 4;  // K runs three blocks ahead of V (S is triple-buffered)
  static constexpr uint32_t kStagesV = 2;
  static constexpr uint32_t kSmemQ = 0;
  static constexpr uint32_t kSmemK = kSmemQ + kTileBytes;
  static constexpr uint32_t kSmemV = kSmemK + kStagesK * kTileBytes;
  // half-row maxima swapped between pair warps: float [2 block parities][2 halves][128 rows]
  static constexpr uint32_t kSmemXmax = kSmemV + kStagesV * kTileBytes;
  static constexpr uint32_t kSmemBar = kSmemXmax + 2 * kHalves * kTileM * 4;
  static constexpr uint32_t kNumBars = 1 + 2 * kStagesK + 2 * kStagesV + 2 * kSBuffers + 2 + 1;
  static constexpr uint32_t kSmemTmemPtr = kSmemBar + kNumBars * 8;
  static constexpr uint32_t kSmemBytes = kSmemTmemPtr + 16;
  // partial row sums for the epilogue; aliases Q, which is dead by then (every S = Q K^T has completed before a
  // softmax thread can leave its loop)
  static constexpr uint32_t kSmemStats = kSmemQ;
  static constexpr uint32_t kTmemS = 0;
  static constexpr uint32_t kTmemO = kSBuffers * kBlockN;
  static constexpr uint32_t kTmemCols = 512;
  static_assert(kTmemO + DPAD <= kTmemCols, "accumulator does not fit TMEM");
  static_assert(kSmemBytes <= 232448, "shared memory over budget");
};

// kTrace: debug instantiation that records clock64() at the pipeline hand-off points of CTA (0,0)
// (scripts/trace_forward.py); the production instantiation compiles all of it away.
constexpr uint32_t kTraceSlots = 8;    // per (role, iteration)
constexpr uint32_t kTraceIters = 128;  // iterations recorded per role
#define MFA_TRACE(role, iter, slot)                                                                   \
  do {                                                                                                \
    if (kTrace && trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 &&             \
        (iter) < kTraceIters)                                                                         \
      trace[((role) * kTraceIters + (iter)) * kTraceSlots + (slot)] = clock64();                      \
  } while (0)

__device__ __forceinline__ void softmax_group_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(kSoftmaxThreads) : "memory");  // the 8 softmax warps only
}
// warp w (columns 0-63) and warp w + 4 (columns 64-127) of the same 32 rows
__device__ __forceinline__ void row_pair_sync(uint32_t quarter) {
  asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
}

template <uint32_t DPAD, bool kBF16, bool kTrace = false>
__global__ void __launch_bounds__(kThreads, 1)
    attention_forward_tcgen05(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                              const __grid_constant__ CUtensorMap mapV, float *__restrict__ O, void *__restrict__ L,
                              uint32_t R, uint32_t C, uint32_t D, float scale_log2, int l_is_fp16,
                              long long *__restrict__ trace) {
  using Cfg = Config<DPAD>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B-swizzled tiles need a 1024 B aligned base

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t head = blockIdx.y;
  const uint32_t q_row0 = blockIdx.x * kTileM;
  const uint32_t num_blocks = (C + kBlockN - 1) / kBlockN;

  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::kSmemBar);
  uint64_t *q_full = bars;
  uint64_t *k_full = q_full + 1;
  uint64_t *k_empty = k_full + Cfg::kStagesK;
  uint64_t *v_full = k_empty + Cfg::kStagesK;
  uint64_t *v_empty = v_full + Cfg::kStagesV;
  uint64_t *s_full = v_empty + Cfg::kStagesV;  // [buffer]  S(i) landed in TMEM
  uint64_t *p_full = s_full + kSBuffers;        // [buffer]  both halves of P(i) written (256 arrivals)
  uint64_t *o_full = p_full + kSBuffers;        // [block parity]  O += P V of a block with that parity is done
  uint64_t *o_final = o_full + 2;               // one-shot: every MMA of this CTA has completed
  float *xmax = reinterpret_cast<float *>(smem + Cfg::kSmemXmax);
  float *stats = reinterpret_cast<float *>(smem + Cfg::kSmemStats);
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(smem + Cfg::kSmemTmemPtr);

  // ---------------- one-time setup ----------------
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (uint32_t s = 0; s < Cfg::kStagesK; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (uint32_t s = 0; s < Cfg::kStagesV; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (uint32_t bf = 0; bf < kSBuffers; ++bf) {
      mbar_init(&s_full[bf], 1);
      mbar_init(&p_full[bf], kSoftmaxThreads);
    }
    mbar_init(&o_full[0], 1);
    mbar_init(&o_full[1], 1);
    mbar_init(o_final, 1);
    fence_barrier_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
    tmem_relinquish();
  }
  if (warp == 9 && lane == 0) {
    prefetch_tensormap(&mapQ);
    prefetch_tensormap(&mapK);
  }
  if (warp == 10 && lane == 0) prefetch_tensormap(&mapV);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < 8) {
    // =====================================================================================
    // softmax warps: thread <-> query row <-> TMEM lane; warpgroup <-> column half
    // =====================================================================================
    setmaxnreg_inc<kSoftmaxRegs>();
    const uint32_t h = warp >> 2;        // column half
    const uint32_t quarter = warp & 3;   // TMEM lane quarter == row-pair id
    const uint32_t row_in_tile = quarter * 32 + lane;
    const uint32_t tLane = tmem_base + ((quarter * 32) << 16);
    const uint32_t tO = tLane + Cfg::kTmemO;
    const uint32_t trace_role = warp == 0 ? 0 : (warp == 4 ? 1 : 3);

    float m = -FLT_MAX;  // running row max (identical in both halves), log2 domain (AttentionKernel+Caching.swift:310)
    float l = 0.f;       // running sum over this half's columns
    // valid columns of this half in the last block (0 when the last block ends before this half starts)
    const uint32_t tail_block = C - (num_blocks - 1) * kBlockN;
    const uint32_t tail_cols = tail_block > h * kHalfN ? min(tail_block - h * kHalfN, kHalfN) : 0u;

    auto s_buffer = [&](uint32_t i) { return tLane + Cfg::kTmemS + (i % kSBuffers) * kBlockN + h * kHalfN; };
    auto load_block = [&](float (&dst)[kHalfN], uint32_t i) {
      mbar_wait(&s_full[i % kSBuffers], (i / kSBuffers) & 1);
      tc_fence_after();
      const uint32_t tS = s_buffer(i);
#pragma unroll
      for (uint32_t c = 0; c < kHalfN; c += 32) tmem_ld32(tS + c, *reinterpret_cast<uint32_t(*)[32]>(&dst[c]));
    };
    // edge mask (maskAttentionMatrixEdge, AttentionKernel+Softmax.swift:228-260) + max of this half of the row
    // (onlineReduceMaximum, :267-287)
    auto half_max = [&](float (&v)[kHalfN], uint32_t i) -> float {
      if (i == num_blocks - 1 && tail_cols < kHalfN) {
#pragma unroll
        for (uint32_t c = 0; c < kHalfN; ++c)
          if (c >= tail_cols) v[c] = -INFINITY;
      }
      float mx0 = v[0], mx1 = v[1], mx2 = v[2], mx3 = v[3];
#pragma unroll
      for (uint32_t c = 4; c < kHalfN; c += 4) {
        mx0 = fmaxf(mx0, v[c]);
        mx1 = fmaxf(mx1, v[c + 1]);
        mx2 = fmaxf(mx2, v[c + 2]);
        mx3 = fmaxf(mx3, v[c + 3]);
      }
      return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
    };
    // publish this half's max of block i / fetch the pair warp's (double-buffered by block parity)
    auto publish_max = [&](float mx, uint32_t i) { xmax[((i & 1) * kHalves + h) * kTileM + row_in_tile] = mx; };
    auto fetch_pair_max = [&](uint32_t i) { return xmax[((i & 1) * kHalves + (h ^ 1)) * kTileM + row_in_tile]; };
    // lazy correction (onlineCorrectO, :290-301): refresh m only when it grew by > 2^8.  `done_blocks` key blocks
    // are already accumulated in O; each half rescales its own DPAD/2 columns of O.
    auto update_max = [&](float row_mx, uint32_t done_blocks) {
      const float m_cand = fmaxf(m, row_mx * scale_log2);
      if (__any_sync(0xffffffffu, m_cand - m > kRescaleThreshold)) {
        if (done_blocks > 0) {
          const float correction = ex2_approx(m - m_cand);
          const uint32_t last = done_blocks - 1;           // O += P V of that block must have landed;
          mbar_wait(&o_full[last & 1], (last >> 1) & 1);   // one barrier per block parity keeps the phase unambiguous
          tc_fence_after();
#pragma unroll
          for (uint32_t c = 0; c < DPAD / 2; c += 32) {
            uint32_t o[32];
            tmem_ld32(tO + h * (DPAD / 2) + c, o);
            tc_wait_ld();
#pragma unroll
            for (uint32_t k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * correction);
            tmem_st32(tO + h * (DPAD / 2) + c, o);
          }
          tc_wait_st();
          l *= correction;
        }
        m = m_cand;
      }
    };
    // P = exp2(S * log2e/sqrt(D) - m) for 32 columns, rounded to the MMA input type and written over S
    // (softmax, :409-416; onlineReduceSum, :304-324)
    auto exp_chunk = [&](const float (&v)[kHalfN], uint32_t c, uint32_t tS, float &sum0, float &sum1) {
      uint32_t packed[16];
#pragma unroll
      for (uint32_t k = 0; k < 16; ++k) {
        const float p0 = ex2_approx(fmaf(v[c + 2 * k], scale_log2, -m));
        const float p1 = ex2_approx(fmaf(v[c + 2 * k + 1], scale_log2, -m));
        sum0 += p0;
        sum1 += p1;
        packed[k] = kBF16 ? pack_bf16x2(p0, p1) : pack_f16x2(p0, p1);
      }
      tmem_st16(tS + (c >> 1), packed);
    };
    // one key block: exp2 stream of block i, with the load / max / pair exchange of block i + 1 folded in
    auto step = [&](float (&cur)[kHalfN], float (&nxt)[kHalfN], uint32_t i) {
      const uint32_t tS = s_buffer(i);
      const bool has_next = i + 1 < num_blocks;
      MFA_TRACE(trace_role, i, 0);
      if (has_next) load_block(nxt, i + 1);  // asynchronous: completes at the tc_wait_ld below
      MFA_TRACE(trace_role, i, 1);
      float sum0 = 0.f, sum1 = 0.f;
      exp_chunk(cur, 0, tS, sum0, sum1);
      float next_mx = 0.f;
      if (has_next) {
        tc_wait_ld();
        next_mx = half_max(nxt, i + 1);  // ALU work the scheduler interleaves with the exp2 stream
        publish_max(next_mx, i + 1);
      }
      MFA_TRACE(trace_role, i, 2);
      exp_chunk(cur, 32, tS, sum0, sum1);
      l += sum0 + sum1;
      MFA_TRACE(trace_role, i, 3);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[i % kSBuffers]);
      MFA_TRACE(trace_role, i, 4);
      if (has_next) {
        row_pair_sync(quarter);
        update_max(fmaxf(next_mx, fetch_pair_max(i + 1)), i + 1);
      }
      MFA_TRACE(trace_role, i, 5);
    };

    {
      float sA[kHalfN], sB[kHalfN];
      load_block(sA, 0);
      tc_wait_ld();
      const float mx0 = half_max(sA, 0);
      publish_max(mx0, 0);
      row_pair_sync(quarter);
      update_max(fmaxf(mx0, fetch_pair_max(0)), 0);
      for (uint32_t i = 0; i < num_blocks; i += 2) {
        step(sA, sB, i);
        if (i + 1 < num_blocks) step(sB, sA, i + 1);
      }
    }

    // ---------------- epilogue: O / l -> global (FP32), L = m + log2(l) ----------------
    stats[h * kTileM + row_in_tile] = l;
    softmax_group_sync();
    const float l_all = l + stats[(h ^ 1) * kTileM + row_in_tile];
    const float inv_l = 1.0f / l_all;

    mbar_wait(o_final, 0);
    tc_fence_after();
    const uint32_t row = q_row0 + row_in_tile;
    float *o_row = O + (static_cast<size_t>(head) * R + row) * D;
    // this warpgroup writes columns [h * DPAD/2, (h+1) * DPAD/2) of O
#pragma unroll
    for (uint32_t cc = 0; cc < DPAD / 2; cc += 32) {
      const uint32_t c = h * (DPAD / 2) + cc;
      uint32_t o[32];
      tmem_ld32(tO + c, o);
      tc_wait_ld();
      if (row < R) {
#pragma unroll
        for (uint32_t k = 0; k < 32; k += 4) {
          if (c + k < D) {  // D % 8 == 0, so a float4 is either fully inside or fully outside
            float4 v = make_float4(__uint_as_float(o[k]) * inv_l, __uint_as_float(o[k + 1]) * inv_l,
                                   __uint_as_float(o[k + 2]) * inv_l, __uint_as_float(o[k + 3]) * inv_l);
            *reinterpret_cast<float4 *>(o_row + c + k) = v;
          }
        }
      }
    }
    if (h == 0 && row < R && L != nullptr) {
      const float lse2 = m + log2f(l_all);  // AttentionKernel+Caching.swift:373-377
      const size_t idx = static_cast<size_t>(head) * R + row;
      if (l_is_fp16)
        reinterpret_cast<__half *>(L)[idx] = __float2half_rn(lse2);
      else
        reinterpret_cast<float *>(L)[idx] = lse2;
    }
  } else {
    setmaxnreg_dec<kOtherRegs>();
    // The producer warps run their control flow warp-wide and hand exactly one elected lane to the
    // TMA / tcgen05 instructions: operands stay in uniform registers and the issue loops are branch-free.
    if (warp == 9) {
      // ===================================================================================
      // TMA producer: Q, then the K stages
      // ===================================================================================
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, Cfg::kTileBytes);
#pragma unroll
        for (uint32_t ds = 0; ds < Cfg::kSubTiles; ++ds)
          tma_load_3d(smem + Cfg::kSmemQ + ds * kSubTileBytes, &mapQ, q_full, ds * 64, q_row0, head);
      }
      for (uint32_t i = 0; i < num_blocks; ++i) {
        const uint32_t stage = i % Cfg::kStagesK, phase = (i / Cfg::kStagesK) & 1;
        mbar_wait(&k_empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[stage], Cfg::kTileBytes);
#pragma unroll
          for (uint32_t ds = 0; ds < Cfg::kSubTiles; ++ds)
            tma_load_3d(smem + Cfg::kSmemK + stage * Cfg::kTileBytes + ds * kSubTileBytes, &mapK, &k_full[stage],
                        ds * 64, i * kBlockN, head);
        }
      }
    } else if (warp == 10) {
      // ===================================================================================
      // TMA producer: the V stages
      // ===================================================================================
      for (uint32_t i = 0; i < num_blocks; ++i) {
        const uint32_t stage = i % Cfg::kStagesV, phase = (i / Cfg::kStagesV) & 1;
        mbar_wait(&v_empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&v_full[stage], Cfg::kTileBytes);
#pragma unroll
          for (uint32_t ds = 0; ds < Cfg::kSubTiles; ++ds)
            tma_load_3d(smem + Cfg::kSmemV + stage * Cfg::kTileBytes + ds * kSubTileBytes, &mapV, &v_full[stage],
                        ds * 64, i * kBlockN, head);
        }
      }
    } else if (warp == 8) {
      // ===================================================================================
      // MMA issuer
      // ===================================================================================
      constexpr uint32_t kFormat = kBF16 ? 1u : 0u;
      // S[128 x 128] = Q[128 x D] . K[128 x D]^T : A and B both K-major
      constexpr uint32_t idescS = make_idesc_f16(kTileM, kBlockN, kFormat, 0, 0);
      // O[128 x DPAD] += P[128 x 128] . V[128 x DPAD] : A from TMEM, B (= V, [key][d]) MN-major
      constexpr uint32_t idescO = make_idesc_f16(kTileM, DPAD, kFormat, 0, 1);
      // Descriptors differ only in the 14-bit start-address field; build each once and add (bytes >> 4).
      const uint64_t descQ = make_smem_desc_sw128(smem_u32(smem + Cfg::kSmemQ), 16, 1024);
      const uint64_t descK = make_smem_desc_sw128(smem_u32(smem + Cfg::kSmemK), 16, 1024);
      const uint64_t descV = make_smem_desc_sw128(smem_u32(smem + Cfg::kSmemV), kSubTileBytes, 1024);

      // every tcgen05.mma / commit below is issued by the one elected lane
      auto issue_S = [&](uint32_t bf, uint32_t stage) {
        const uint32_t d_tmem = tmem_base + Cfg::kTmemS + bf * kBlockN;
        const uint64_t b0 = descK + ((stage * Cfg::kTileBytes) >> 4);
#pragma unroll
        for (uint32_t k = 0; k < DPAD / 16; ++k) {
          // 16 elements = 32 B inside the 128 B swizzle row; 4 k-steps per 64-element sub-tile
          const uint32_t off = ((k >> 2) * kSubTileBytes + (k & 3) * 32) >> 4;
          umma_ss(d_tmem, descQ + off, b0 + off, idescS, k > 0);
        }
      };
      auto issue_PV = [&](uint32_t bf, uint32_t stage, uint32_t accumulate) {
        const uint32_t d_tmem = tmem_base + Cfg::kTmemO;
        const uint64_t b0 = descV + ((stage * Cfg::kTileBytes) >> 4);
#pragma unroll
        for (uint32_t k = 0; k < kBlockN / 16; ++k) {
          // P of keys [64 hh, 64 hh + 64) sits in columns [64 hh, 64 hh + 32) of the S buffer (two 16-bit keys per
          // column); V: 16 keys = two 8-row groups of 1024 B
          const uint32_t a_tmem = tmem_base + Cfg::kTmemS + bf * kBlockN + (k >> 2) * kHalfN + (k & 3) * 8;
          umma_ts(d_tmem, a_tmem, b0 + ((k * 2048) >> 4), idescO, k > 0 ? 1u : accumulate);
        }
      };

      // prologue: S(0) .. S(2)
      mbar_wait(q_full, 0);
      for (uint32_t i = 0; i < kSBuffers && i < num_blocks; ++i) {
        mbar_wait(&k_full[i], 0);
        tc_fence_after();
        if (elect_one()) {
          issue_S(i, i);
          umma_commit(&s_full[i]);
          umma_commit(&k_empty[i]);
        }
        __syncwarp();
      }

      for (uint32_t i = 0; i < num_blocks; ++i) {
        const uint32_t bf = i % kSBuffers, ph = (i / kSBuffers) & 1;
        const uint32_t stage = i % Cfg::kStagesV, phase = (i / Cfg::kStagesV) & 1;
        const uint32_t ni = i + kSBuffers;  // the S block that reuses this buffer
        const uint32_t nstage = ni % Cfg::kStagesK, nphase = (ni / Cfg::kStagesK) & 1;
        const bool has_next = ni < num_blocks;
        const bool last = i + 1 == num_blocks;
        mbar_wait(&v_full[stage], phase);
        MFA_TRACE(2, i, 0);
        mbar_wait(&p_full[bf], ph);
        tc_fence_after();
        MFA_TRACE(2, i, 1);
        if (elect_one()) {
          issue_PV(bf, stage, i > 0 ? 1u : 0u);
          umma_commit(&o_full[i & 1]);
          umma_commit(&v_empty[stage]);
          if (last) umma_commit(o_final);
        }
        __syncwarp();
        MFA_TRACE(2, i, 2);
        if (has_next) {
          mbar_wait(&k_full[nstage], nphase);
          tc_fence_after();
          MFA_TRACE(2, i, 3);
          if (elect_one()) {
            issue_S(bf, nstage);  // overwrites P(i) only after P V(i): the tensor pipe runs in order
            umma_commit(&s_full[bf]);
            umma_commit(&k_empty[nstage]);
          }
          __syncwarp();
        }
        MFA_TRACE(2, i, 4);
      }
    }
  }

  // ---------------- teardown ----------------
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <uint32_t DPAD, bool kBF16, bool kTrace = false>
cudaError_t launch(const AttentionParams &p, cudaStream_t stream, long long *trace = nullptr) {
  using Cfg = Config<DPAD>;
  auto kernel = attention_forward_tcgen05<DPAD, kBF16, kTrace>;
  static std::once_flag once;
  static cudaError_t attr_status = cudaSuccess;
  std::call_once(once, [&] {
    attr_status = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  });
  if (attr_status != cudaSuccess) return attr_status;

  CUtensorMap mapQ, mapK, mapV;
  cudaError_t e;
  if ((e = make_tensor_map_16bit(&mapQ, p.buf[sQ], p.R, p.D, p.batch, kTileM)) != cudaSuccess) return e;
  if ((e = make_tensor_map_16bit(&mapK, p.buf[sK], p.C, p.D, p.batch, kBlockN)) != cudaSuccess) return e;
  if ((e = make_tensor_map_16bit(&mapV, p.buf[sV], p.C, p.D, p.batch, kBlockN)) != cudaSuccess) return e;

  dim3 grid((p.R + kTileM - 1) / kTileM, p.batch);
  kernel<<<grid, kThreads, Cfg::kSmemBytes, stream>>>(mapQ, mapK, mapV, static_cast<float *>(p.buf[sO]), p.buf[sL],
                                                      p.R, p.C, p.D, p.scale_log2, p.prec[sL] == FP16 ? 1 : 0, trace);
  return cudaGetLastError();
}

}  // namespace fwd

uint32_t tcgen05_forward_max_head() { return 128; }

bool tcgen05_forward_supported(const AttentionParams &p) {
  return (p.prec[sQ] == FP16 || p.prec[sQ] == BF16) && p.prec[sK] == p.prec[sQ] && p.prec[sV] == p.prec[sQ] &&
         p.prec[sO] == FP32 && p.D % 8 == 0 && p.D <= tcgen05_forward_max_head() && !p.transposed[sQ] &&
         !p.transposed[sK] && !p.transposed[sV] && !p.transposed[sO];
}

cudaError_t launch_tcgen05_forward(const AttentionParams &p, cudaStream_t stream) {
  if (!tcgen05_forward_supported(p)) {
    set_launch_detail("descriptor is outside the tcgen05 forward kernel's domain");
    return cudaErrorInvalidValue;
  }
  const bool bf16 = p.prec[sQ] == BF16;
  if (p.D <= 64) return bf16 ? fwd::launch<64, true>(p, stream) : fwd::launch<64, false>(p, stream);
  return bf16 ? fwd::launch<128, true>(p, stream) : fwd::launch<128, false>(p, stream);
}

void set_forward_stagger(uint32_t) {}  // no tunable left in this kernel generation

// Debug entry (not in include/mfa_b200.h): the D=128 bf16 forward with pipeline timestamps of CTA (0,0)
// written to `trace` (4 roles x 128 iterations x 8 slots of clock64()).  Used by scripts/trace_forward.py.
cudaError_t launch_tcgen05_forward_trace(const AttentionParams &p, cudaStream_t stream, long long *trace) {
  return fwd::launch<128, true, true>(p, stream, trace);
}

void tcgen05_forward_geometry(uint32_t D, uint32_t *threads, uint32_t *smem_bytes, uint32_t *par, uint32_t *trav,
                              uint32_t *head) {
  *threads = fwd::kThreads;
  *smem_bytes = D <= 64 ? fwd::Config<64>::kSmemBytes : fwd::Config<128>::kSmemBytes;
  *par = fwd::kTileM;
  *trav = fwd::kBlockN;
  *head = D <= 64 ? 64 : 128;
  const uint32_t padded = (D + 7) / 8 * 8;
  if (*head > padded) *head = padded;
}

// ---- backward tcgen05 kernels are not built yet: the heuristic never selects them --------------
uint32_t tcgen05_backward_max_head() { return 0; }
bool tcgen05_backward_supported(const AttentionParams &) { return false; }
cudaError_t launch_tcgen05_backward_query(const AttentionParams &, cudaStream_t) {
  set_launch_detail("tcgen05 backward-query kernel is not compiled in");
  return cudaErrorNotSupported;
}
cudaError_t launch_tcgen05_backward_key_value(const AttentionParams &, cudaStream_t) {
  set_launch_detail("tcgen05 backward-key-value kernel is not compiled in");
  return cudaErrorNotSupported;
}
void tcgen05_backward_geometry(int, uint32_t, uint32_t *threads, uint32_t *smem_bytes, uint32_t *par, uint32_t *trav,
                               uint32_t *head) {
  *threads = *smem_bytes = *par = *trav = *head = 0;
}

}  // namespace mfa
