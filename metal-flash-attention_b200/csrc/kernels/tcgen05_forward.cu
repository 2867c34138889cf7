// FlashAttention forward for sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM.
//
// Replaces the reference's generated forward kernel (loopForward, AttentionKernel+Source.swift:158-200;
// outer product S = Q K^T, +OuterProduct.swift:18-487; online softmax, +Softmax.swift:228-324,334-505;
// accumulate O += P V, +Accumulate.swift:24-582) for 16-bit row-major operands.
//
// One CTA owns one 128-row tcgen05 M-tile of Q and walks the keys in blocks of 128.  Inside the CTA the key
// axis of every block is split in two halves that are treated as two independent attention streams ("lo" =
// keys 0-63 of each block, "hi" = keys 64-127): each half has its own softmax warpgroup, its own running
// (max, sum) and its own O accumulator in TMEM, and the two partial results are merged once, in the epilogue
// (the standard split-KV combine).  That removes every per-block exchange between the two warpgroups while
// letting 256 threads share one row block.  Warp roles (384 threads):
//   warps 0-3   softmax for the lo key half (thread = one query row = one TMEM lane)
//   warps 4-7   softmax for the hi key half (same rows, the other 64 columns of S)
//   warp  8     MMA issuer (one elected lane issues every tcgen05.mma / commit); owns the TMEM allocation
//   warp  9     TMA producer (Q once, then K and V stages)
//   warps 10-11 idle (they donate their registers via setmaxnreg)
// On-chip residency (the reference's "cache Q, O" rows, AttentionDescriptor+Parameters.swift:109-120,
// re-expressed for B200): Q stays in SMEM for the whole traversal, both O accumulators stay in TMEM; S is
// double-buffered in TMEM and each half is overwritten in place by its P (16-bit), which feeds the second
// MMA straight from TMEM.  TMEM columns: [0,128) S/P buffer 0, [128,256) S/P buffer 1, [256,256+D) O_lo,
// [256+D,256+2D) O_hi.  S(i+1) is computed while the softmax warps work on S(i), so the tensor pipe
// (S: 8 MMAs, P V: 2 x 4 MMAs per block) and the MUFU pipe (128 x 128 exp2 per block) overlap.
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
constexpr uint32_t kHalfN = 64;    // keys per softmax warpgroup per block
constexpr uint32_t kHalves = 2;
constexpr uint32_t kSBuffers = 2;  // S/P buffers
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
  static constexpr uint32_t kStages = 3;
  static constexpr uint32_t kSmemQ = 0;
  static constexpr uint32_t kSmemK = kSmemQ + kTileBytes;
  static constexpr uint32_t kSmemV = kSmemK + kStages * kTileBytes;
  // float2 (m, l) [2 halves][128 rows] for the epilogue merge; aliases Q, which is dead by then (every
  // S = Q K^T has completed before a softmax thread can leave its loop)
  static constexpr uint32_t kSmemStats = kSmemQ;
  static constexpr uint32_t kSmemBar = kSmemV + kStages * kTileBytes;
  static constexpr uint32_t kNumBars = 1 + 4 * kStages + kSBuffers + kHalves * kSBuffers + kHalves + 1;
  static constexpr uint32_t kSmemTmemPtr = kSmemBar + kNumBars * 8;
  static constexpr uint32_t kSmemBytes = kSmemTmemPtr + 16 + 1024;  // + slack for manual 1024 B alignment
  static constexpr uint32_t kTmemS = 0;
  static constexpr uint32_t kTmemO = kSBuffers * kBlockN;
  static constexpr uint32_t kTmemCols = 512;
  static_assert(kTmemO + kHalves * DPAD <= kTmemCols, "accumulators do not fit TMEM");
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

template <uint32_t DPAD, bool kBF16, bool kTrace = false>
__global__ void __launch_bounds__(kThreads, 1)
    attention_forward_tcgen05(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                              const __grid_constant__ CUtensorMap mapV, float *__restrict__ O, void *__restrict__ L,
                              uint32_t R, uint32_t C, uint32_t D, float scale_log2, int l_is_fp16,
                              uint32_t stagger_cycles, long long *__restrict__ trace) {
  using Cfg = Config<DPAD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t head = blockIdx.y;
  const uint32_t q_row0 = blockIdx.x * kTileM;
  const uint32_t num_blocks = (C + kBlockN - 1) / kBlockN;

  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::kSmemBar);
  uint64_t *q_full = bars;
  uint64_t *k_full = q_full + 1;
  uint64_t *k_empty = k_full + Cfg::kStages;
  uint64_t *v_full = k_empty + Cfg::kStages;
  uint64_t *v_empty = v_full + Cfg::kStages;
  uint64_t *s_full = v_empty + Cfg::kStages;       // [buffer]        S(i) landed in TMEM
  uint64_t *p_full = s_full + kSBuffers;            // [half][buffer]  P half written by its warpgroup
  uint64_t *o_full = p_full + kHalves * kSBuffers;  // [half]          one phase per key block: O_half += P V done
  uint64_t *o_final = o_full + kHalves;             // one-shot: every MMA of this CTA has completed
  float2 *stats = reinterpret_cast<float2 *>(smem + Cfg::kSmemStats);
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(smem + Cfg::kSmemTmemPtr);

  // ---------------- one-time setup ----------------
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (uint32_t s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (uint32_t bf = 0; bf < kSBuffers; ++bf) {
      mbar_init(&s_full[bf], 1);
      for (uint32_t h = 0; h < kHalves; ++h) mbar_init(&p_full[h * kSBuffers + bf], kTileM);
    }
    for (uint32_t h = 0; h < kHalves; ++h) mbar_init(&o_full[h], 1);
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
    prefetch_tensormap(&mapV);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < 8) {
    // =====================================================================================
    // softmax warps: thread <-> query row <-> TMEM lane; warpgroup <-> key half
    // =====================================================================================
    setmaxnreg_inc<kSoftmaxRegs>();
    const uint32_t h = warp >> 2;  // key half
    const uint32_t row_in_tile = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = ((warp & 3) * 32) << 16;  // this warp's TMEM lane quarter
    const uint32_t tLane = tmem_base + lane_addr;
    const uint32_t tO = tLane + Cfg::kTmemO + h * DPAD;
    const uint32_t trace_role = warp == 0 ? 0 : (warp == 4 ? 1 : 3);

    float m = -FLT_MAX;  // running max of this half, log2 domain   (AttentionKernel+Caching.swift:310)
    float l = 0.f;       // running sum of this half
    // valid columns of this half in the last block (0 when the last block ends before this half starts)
    const uint32_t tail_block = C - (num_blocks - 1) * kBlockN;
    const uint32_t tail_cols = tail_block > h * kHalfN ? min(tail_block - h * kHalfN, kHalfN) : 0u;

    // Software pipeline over key blocks: while the exp2 stream of block i occupies the MUFU pipe, the same
    // thread already loads S(i+1) from TMEM and reduces its row max on the ALU pipe, so the MUFU pipe never
    // idles between blocks.  `cur` holds S(i) (its max is already folded into m), `nxt` receives S(i+1).
    auto load_block = [&](float (&dst)[kHalfN], uint32_t i) {
      mbar_wait(&s_full[i & 1], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t tS = tLane + Cfg::kTmemS + (i & 1) * kBlockN + h * kHalfN;
#pragma unroll
      for (uint32_t c = 0; c < kHalfN; c += 32) tmem_ld32(tS + c, *reinterpret_cast<uint32_t(*)[32]>(&dst[c]));
    };
    // edge mask (maskAttentionMatrixEdge, AttentionKernel+Softmax.swift:228-260) + row max of one block
    // (onlineReduceMaximum, :267-287): this half of the row sits in this thread's registers
    auto block_max = [&](float (&v)[kHalfN], uint32_t i) -> float {
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
    // lazy correction (onlineCorrectO, :290-301): refresh m only when it grew by > 2^8.  `done_blocks` key
    // blocks have already been accumulated into O_half and must be rescaled.
    auto update_max = [&](float block_mx, uint32_t done_blocks) {
      const float m_cand = fmaxf(m, block_mx * scale_log2);
      if (__any_sync(0xffffffffu, m_cand - m > kRescaleThreshold)) {
        if (done_blocks > 0) {
          const float correction = ex2_approx(m - m_cand);
          mbar_wait(&o_full[h], (done_blocks - 1) & 1);  // O_half += P V of the previous block has landed
          tc_fence_after();
#pragma unroll
          for (uint32_t c = 0; c < DPAD; c += 32) {
            uint32_t o[32];
            tmem_ld32(tO + c, o);
            tc_wait_ld();
#pragma unroll
            for (uint32_t k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * correction);
            tmem_st32(tO + c, o);
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
    auto step = [&](float (&cur)[kHalfN], float (&nxt)[kHalfN], uint32_t i) {
      const uint32_t bf = i & 1;
      const uint32_t tS = tLane + Cfg::kTmemS + bf * kBlockN + h * kHalfN;
      const bool has_next = i + 1 < num_blocks;
      MFA_TRACE(trace_role, i, 0);
      if (has_next) load_block(nxt, i + 1);  // asynchronous: completes at the tc_wait_ld below
      MFA_TRACE(trace_role, i, 1);
      float sum0 = 0.f, sum1 = 0.f;
      exp_chunk(cur, 0, tS, sum0, sum1);
      float next_mx = 0.f;
      if (has_next) {
        tc_wait_ld();
        next_mx = block_max(nxt, i + 1);  // ALU work the scheduler interleaves with the exp2 stream below
      }
      MFA_TRACE(trace_role, i, 2);
      exp_chunk(cur, 32, tS, sum0, sum1);
      l += sum0 + sum1;
      MFA_TRACE(trace_role, i, 3);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[h * kSBuffers + bf]);
      MFA_TRACE(trace_role, i, 4);
      if (has_next) update_max(next_mx, i + 1);
    };

    {
      float sA[kHalfN], sB[kHalfN];
      load_block(sA, 0);
      tc_wait_ld();
      update_max(block_max(sA, 0), 0);
      for (uint32_t i = 0; i < num_blocks; i += 2) {
        step(sA, sB, i);
        if (i + 1 < num_blocks) step(sB, sA, i + 1);
      }
    }


    // ---------------- epilogue: merge the two key halves, O / l -> global (FP32), L = m + log2(l) --------
    stats[h * kTileM + row_in_tile] = make_float2(m, l);
    softmax_group_sync();
    const float2 other = stats[(h ^ 1) * kTileM + row_in_tile];
    const float m_all = fmaxf(m, other.x);
    const float a_mine = ex2_approx(m - m_all), a_other = ex2_approx(other.x - m_all);
    const float l_all = fmaf(l, a_mine, other.y * a_other);
    const float inv_l = 1.0f / l_all;
    const float w_lo = (h == 0 ? a_mine : a_other) * inv_l, w_hi = (h == 0 ? a_other : a_mine) * inv_l;

    mbar_wait(o_final, 0);
    tc_fence_after();
    const uint32_t row = q_row0 + row_in_tile;
    float *o_row = O + (static_cast<size_t>(head) * R + row) * D;
    // this warpgroup writes columns [h * DPAD/2, (h+1) * DPAD/2) of the merged O
    constexpr uint32_t kColsPerGroup = DPAD / 2;
#pragma unroll
    for (uint32_t cc = 0; cc < kColsPerGroup; cc += 32) {
      const uint32_t c = h * kColsPerGroup + cc;
      uint32_t lo[32], hi[32];
      tmem_ld32(tLane + Cfg::kTmemO + c, lo);
      tmem_ld32(tLane + Cfg::kTmemO + DPAD + c, hi);
      tc_wait_ld();
      if (row < R) {
#pragma unroll
        for (uint32_t k = 0; k < 32; k += 4) {
          if (c + k < D) {  // D % 8 == 0, so a float4 is either fully inside or fully outside
            float4 v;
            v.x = fmaf(__uint_as_float(lo[k]), w_lo, __uint_as_float(hi[k]) * w_hi);
            v.y = fmaf(__uint_as_float(lo[k + 1]), w_lo, __uint_as_float(hi[k + 1]) * w_hi);
            v.z = fmaf(__uint_as_float(lo[k + 2]), w_lo, __uint_as_float(hi[k + 2]) * w_hi);
            v.w = fmaf(__uint_as_float(lo[k + 3]), w_lo, __uint_as_float(hi[k + 3]) * w_hi);
            *reinterpret_cast<float4 *>(o_row + c + k) = v;
          }
        }
      }
    }
    if (h == 0 && row < R && L != nullptr) {
      const float lse2 = m_all + log2f(l_all);  // AttentionKernel+Caching.swift:373-377
      const size_t idx = static_cast<size_t>(head) * R + row;
      if (l_is_fp16)
        reinterpret_cast<__half *>(L)[idx] = __float2half_rn(lse2);
      else
        reinterpret_cast<float *>(L)[idx] = lse2;
    }
  } else {
    setmaxnreg_dec<kOtherRegs>();
    // Both producer warps run their control flow warp-wide and hand exactly one elected lane to the
    // TMA / tcgen05 instructions: operands stay in uniform registers and the issue loops are branch-free.
    if (warp == 9) {
      // ===================================================================================
      // TMA producer
      // ===================================================================================
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, Cfg::kTileBytes);
#pragma unroll
        for (uint32_t ds = 0; ds < Cfg::kSubTiles; ++ds)
          tma_load_3d(smem + Cfg::kSmemQ + ds * kSubTileBytes, &mapQ, q_full, ds * 64, q_row0, head);
      }
      for (uint32_t i = 0; i < num_blocks; ++i) {
        const uint32_t stage = i % Cfg::kStages, phase = (i / Cfg::kStages) & 1;
        mbar_wait(&k_empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[stage], Cfg::kTileBytes);
#pragma unroll
          for (uint32_t ds = 0; ds < Cfg::kSubTiles; ++ds)
            tma_load_3d(smem + Cfg::kSmemK + stage * Cfg::kTileBytes + ds * kSubTileBytes, &mapK, &k_full[stage],
                        ds * 64, i * kBlockN, head);
        }
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
      // O_half[128 x DPAD] += P_half[128 x 64] . V_half[64 x DPAD] : A from TMEM, B (= V, [key][d]) MN-major
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
      auto issue_PV = [&](uint32_t h, uint32_t bf, uint32_t stage, uint32_t accumulate) {
        const uint32_t d_tmem = tmem_base + Cfg::kTmemO + h * DPAD;
        const uint32_t a_tmem = tmem_base + Cfg::kTmemS + bf * kBlockN + h * kHalfN;
        // keys [64 h, 64 h + 64) of the stage: 16 keys = two 8-row groups of 1024 B
        const uint64_t b0 = descV + ((stage * Cfg::kTileBytes + h * (kHalfN / 16) * 2048) >> 4);
#pragma unroll
        for (uint32_t k = 0; k < kHalfN / 16; ++k)
          umma_ts(d_tmem, a_tmem + k * 8, b0 + ((k * 2048) >> 4), idescO, k > 0 ? 1u : accumulate);
      };

      // prologue: S(0) and S(1)
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
        const uint32_t bf = i & 1, ph = (i >> 1) & 1;
        const uint32_t stage = i % Cfg::kStages, phase = (i / Cfg::kStages) & 1;
        const uint32_t ni = i + kSBuffers;  // the S block that reuses this buffer
        const uint32_t nstage = ni % Cfg::kStages, nphase = (ni / Cfg::kStages) & 1;
        const bool has_next = ni < num_blocks;
        const bool last = i + 1 == num_blocks;
        mbar_wait(&v_full[stage], phase);
        MFA_TRACE(2, i, 0);
#pragma unroll
        for (uint32_t h = 0; h < kHalves; ++h) {
          mbar_wait(&p_full[h * kSBuffers + bf], ph);
          tc_fence_after();
          MFA_TRACE(2, i, 1 + 2 * h);
          if (elect_one()) {
            issue_PV(h, bf, stage, i > 0 ? 1u : 0u);
            umma_commit(&o_full[h]);
            if (h == kHalves - 1) {
              umma_commit(&v_empty[stage]);
              if (last) umma_commit(o_final);
            }
          }
          __syncwarp();
          MFA_TRACE(2, i, 2 + 2 * h);
        }
        if (has_next) {
          mbar_wait(&k_full[nstage], nphase);
          tc_fence_after();
          if (elect_one()) {
            issue_S(bf, nstage);  // overwrites P(i) only after both P V(i): the tensor pipe runs in order
            umma_commit(&s_full[bf]);
            umma_commit(&k_empty[nstage]);
          }
          __syncwarp();
        }
        MFA_TRACE(2, i, 5);
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

uint32_t g_stagger_cycles = 600;  // tuned on B200 (scripts/tune_forward.py); settable through the debug hook

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
                                                      p.R, p.C, p.D, p.scale_log2, p.prec[sL] == FP16 ? 1 : 0,
                                                      g_stagger_cycles, trace);
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

void set_forward_stagger(uint32_t cycles) { fwd::g_stagger_cycles = cycles; }

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
