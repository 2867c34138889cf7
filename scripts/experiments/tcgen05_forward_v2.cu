// FlashAttention forward for sm_100a, head dimensions D <= 128, 16-bit row-major operands: TMA -> shared memory ->
// tcgen05.mma -> TMEM, persistent CTAs.
//
// Replaces the reference's generated forward kernel (loopForward, AttentionKernel+Source.swift:158-200; outer product
// S = Q K^T, +OuterProduct.swift:18-487; online softmax, +Softmax.swift:228-324,334-505; accumulate O += P V,
// +Accumulate.swift:24-582).
//
// Why this shape.  The exponentials are the co-bottleneck of the forward pass: 128 x 128 of them per tile-block need
// 1024 MUFU cycles (16 ex2 / clk / SM), exactly what the tile-block's two GEMMs need on the tensor pipe at D = 128.  The
// exp-stream micro-benchmark (tests/gpu_probe/exp_probe.cu, profiles/r2_exp_probe.txt) shows that a warp that issues an
// ex2 cannot issue anything else for 8 cycles and that one warp per SM sub-partition reaches 12.5 exp / clk / SM on the
// softmax loop, two warps 14.3.  The first version of this kernel gave each of two ping-pong tiles its own softmax
// warpgroup (one thread per row, 128 columns): every tile's step then ran at the ONE-warp-per-sub-partition rate
// (1400-2100 cycles) and sat in the tile's dependency chain S -> softmax -> P V -> next S, 3400 cycles per block
// pair for 2048 of tensor work (58 % tensor-pipe utilisation, 0.76 of the measured cuBLAS peak).  Here ONE 128-row
// tile is processed by BOTH softmax warpgroups (two threads per row, 64 columns each -- the organisation of the
// D <= 256 kernel), so a tile-block's exponentials take ~1100-1300 cycles, and S is double- (or triple-) buffered in
// TMEM so that S(i+1) is computed while the warps work on S(i): the softmax warps never wait for the tensor pipe in
// steady state and the pipe only waits for P.
//
// Warp roles (384 threads):
//   warps 0-7   softmax: thread <-> (query row, 64-column half of the block); warps w and w + 4 share TMEM lane
//               quarter w and take the (rare) rescale decision jointly through a 64-thread named barrier
//   warp  8     MMA issuer (one elected thread issues every tcgen05.mma / commit); owns TMEM alloc
//   warp  9     TMA producer: Q (double-buffered across work items) and the K ring
//   warp  10    TMA producer: the V ring
//   warp  11    idle (donates registers)
// Persistent CTAs walk work items (head, 128-row tile[, key split]); barrier phases are carried across items, so the
// producers run ahead into the next item (Q prefetch, first S tiles) while the softmax warps store the current one.
//   TMEM columns: S/P buffers [0, 128 kSBuffers), then O [.., + DPAD)
// Softmax bookkeeping follows Appendix A of SURVEY.md (log2 domain, L = m + log2 l); the running max is refreshed
// lazily (only when a half-row of P could exceed 2^8), which is mathematically identical.
// Split-KV for small grids: see the launcher.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "attention_params.h"
#include "device_state.h"
#include "sm100_ptx.cuh"
#include "tma_host.h"

namespace mfa {

// Split-KV has two forms (tests cover both; mfa_debug_set_forward_fused(0) forces the second):
//   fused   -- ONE launch: every split CTA leaves its raw partial (unnormalised O, m, l) in the library's workspace,
//              announces it on a per-tile arrival counter, waits for its siblings and then merges and stores a
//              1 / num_splits slice of the tile's rows (all CTAs are co-resident: cooperative launch, one item each);
//   scratch -- TWO launches: normalised partials + the combine_splits kernel (the fallback when a cooperative launch
//              is refused, e.g. on a partitioned GPU).
static int g_forward_fused_enabled = 1;
void tcgen05_forward_set_fused(int enabled) { g_forward_fused_enabled = enabled; }

namespace fwd2 {

using namespace ptx;

constexpr uint32_t kTileM = 128;               // query rows per work item (one tcgen05 M-tile)
constexpr uint32_t kBlockN = 128;              // keys per traversal block
constexpr uint32_t kCols = kBlockN / 2;        // columns of a block per softmax thread
constexpr uint32_t kSubTileBytes = 128 * 128;  // [128 rows][64 x 16-bit] = one 128B-swizzled TMA box
constexpr uint32_t kThreads = 384;
constexpr uint32_t kSoftmaxThreads = 256;
// setmaxnreg budget: launched with floor(65536 / 384 / 8) * 8 = 168 registers per thread; the softmax warpgroups grow
// after the producer warpgroup has shrunk.  The sum must not exceed the launch allocation.
constexpr uint32_t kLaunchRegs = 168, kSoftmaxRegs = 208, kOtherRegs = 88;
static_assert(kSoftmaxRegs * 256 + kOtherRegs * 128 <= kLaunchRegs * kThreads, "setmaxnreg over-subscribed");
constexpr float kLazySumLimit = 256.0f;  // a half-row of P summing to <= 2^8 proves every element is <= 2^8

// ---- table-driven tuning parameters (AttentionDescriptor+Parameters analogue; see descriptor.cpp) ----------------
// kSBuffers: S/P buffers in TMEM (2 or 3).  kPoly: of every 4 element pairs, how many take exp2 on the FMA pipe.
#ifndef MFA_FWD2_SBUFFERS
#define MFA_FWD2_SBUFFERS 2
#endif
#ifndef MFA_FWD2_POLY_D128
#define MFA_FWD2_POLY_D128 0
#endif
#ifndef MFA_FWD2_POLY_D64
#define MFA_FWD2_POLY_D64 1
#endif

template <uint32_t DPAD, uint32_t kSBuffers>
struct Config {
  static constexpr uint32_t kSubTiles = DPAD / 64;                   // 64-element sub-tiles along D
  static constexpr uint32_t kTileBytes = kSubTiles * kSubTileBytes;  // one 128 x DPAD operand tile
  static constexpr uint32_t kStages = DPAD <= 64 ? 4 : 2;            // K ring and V ring
  static constexpr uint32_t kSmemQ = 0;                              // two buffers: the next item's Q is prefetched
  static constexpr uint32_t kSmemK = kSmemQ + 2 * kTileBytes;
  static constexpr uint32_t kSmemV = kSmemK + kStages * kTileBytes;
  static constexpr uint32_t kSmemScratch = kSmemV + kStages * kTileBytes;  // epilogue transpose: 8 warps x 4 KB
  static constexpr uint32_t kSmemXch = kSmemScratch + 8 * 4096;           // float [2][128]: row max / row sum exchange
  static constexpr uint32_t kSmemBar = kSmemXch + 2 * kTileM * 4;
  // q_full[2] q_empty[2] k_full[st] k_empty[st] v_full[st] v_empty[st] s_full[sb] p_full[sb][2] o_full o_final o_free
  static constexpr uint32_t kNumBars = 4 + 4 * kStages + 3 * kSBuffers + 3;
  static constexpr uint32_t kSmemTmemPtr = kSmemBar + kNumBars * 8;
  static constexpr uint32_t kSmemBytes = kSmemTmemPtr + 16;
  static_assert(kSmemBytes <= 232448, "shared memory over budget");
  static constexpr uint32_t kTmemO = kSBuffers * kBlockN;
  static constexpr uint32_t kTmemCols = 512;
  static_assert(kTmemO + DPAD <= kTmemCols, "tile does not fit TMEM");
};

// named barrier over `count` threads with an OR reduction of `flag` (all participating warps get the result)
__device__ __forceinline__ bool bar_red_or(uint32_t id, uint32_t count, bool flag) {
  uint32_t out;
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.u32 q, %3, 0;\n"
      "barrier.cta.red.or.pred p, %1, %2, q;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(out)
      : "r"(id), "r"(count), "r"(static_cast<uint32_t>(flag))
      : "memory");
  return out != 0;
}
__device__ __forceinline__ void bar_sync(uint32_t id, uint32_t count) {
  asm volatile("barrier.cta.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// kTrace: debug instantiation that records clock64() at the pipeline hand-off points of CTA 0
// (scripts/trace_forward.py); the production instantiation compiles all of it away.
constexpr uint32_t kTraceSlots = 8;   // per (role, iteration)
constexpr uint32_t kTraceIters = 64;  // iterations recorded per role
#define MFA_TRACE(role, iter, slot)                                                                    \
  do {                                                                                                 \
    if (kTrace && trace != nullptr && blockIdx.x == 0 && lane == 0 && (iter) < kTraceIters)            \
      trace[((role) * kTraceIters + (iter)) * kTraceSlots + (slot)] = clock64();                        \
  } while (0)

template <uint32_t DPAD, bool kBF16, uint32_t kSBuffers, uint32_t kPoly, bool kFused, bool kTrace = false>
__global__ void __launch_bounds__(kThreads, 1)
    attention_forward_tcgen05(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                              const __grid_constant__ CUtensorMap mapV, float *__restrict__ O, void *__restrict__ L,
                              uint32_t R, uint32_t C, uint32_t D, float scale_log2, int l_is_fp16,
                              uint32_t num_items, uint32_t tiles_per_head, uint32_t num_splits, uint32_t batch,
                              float *__restrict__ part_O, float2 *__restrict__ part_ml,
                              uint32_t *__restrict__ counters, long long *__restrict__ trace) {
  using Cfg = Config<DPAD, kSBuffers>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Split-KV: when there are fewer (head, tile) items than SMs, the key axis of every item is cut into num_splits
  // equal ranges that become separate work items (see the launcher for the two ways the partials are merged).
  const uint32_t total_blocks = (C + kBlockN - 1) / kBlockN;
  const uint32_t num_blocks = total_blocks / num_splits;  // key blocks per work item (host guarantees divisibility)

  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::kSmemBar);
  uint64_t *q_full = bars;                          // [2]
  uint64_t *q_empty = q_full + 2;                   // [2] every S MMA of the item has read this Q buffer
  uint64_t *k_full = q_empty + 2;                   // [stages]
  uint64_t *k_empty = k_full + Cfg::kStages;
  uint64_t *v_full = k_empty + Cfg::kStages;
  uint64_t *v_empty = v_full + Cfg::kStages;
  uint64_t *s_full = v_empty + Cfg::kStages;        // [S buffer]
  uint64_t *p_full = s_full + kSBuffers;            // [S buffer][column half] (128 arrivals each)
  uint64_t *o_full = p_full + 2 * kSBuffers;        // one phase per key block: O += P V of the block has landed
  uint64_t *o_final = o_full + 1;                   // one phase per item: the item's last O += P V has landed
  uint64_t *o_free = o_final + 1;                   // one phase per item: the epilogue has read O out of TMEM (256)
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(smem + Cfg::kSmemTmemPtr);

  // ---------------- one-time setup ----------------
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
    }
    for (uint32_t s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (uint32_t bf = 0; bf < kSBuffers; ++bf) {
      mbar_init(&s_full[bf], 1);
      mbar_init(&p_full[2 * bf], kTileM);
      mbar_init(&p_full[2 * bf + 1], kTileM);
    }
    mbar_init(o_full, 1);
    mbar_init(o_final, 1);
    mbar_init(o_free, kSoftmaxThreads);
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
  // scratch split-KV: let the combine kernel (launched with programmatic stream serialisation) be set up now; its
  // griddepcontrol.wait still holds it until this grid has completed and flushed
  if (!kFused && num_splits > 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp < 8) {
    // =====================================================================================
    // softmax warps: thread <-> (query row, half of the block's 128 key columns).  Warpgroup h = warp / 4 owns columns
    // [64 h, 64 h + 64); warps w and w + 4 share the 32 rows of TMEM lane quarter w and take their joint decisions
    // (lazy rescale) through a 64-thread named barrier with an OR reduction.
    // =====================================================================================
    setmaxnreg_inc<kSoftmaxRegs>();
    const uint32_t h = warp >> 2, quarter = warp & 3;
    const uint32_t row_in_tile = quarter * 32 + lane;
    const uint32_t tTile = tmem_base + ((quarter * 32) << 16);
    const uint32_t tO = tTile + Cfg::kTmemO;
    const uint32_t pair_bar = 2 + quarter;  // named barrier of the two warps that share these rows
    const uint32_t trace_role = warp == 0 ? 0 : (warp == 4 ? 1 : 7);
    float *xch = reinterpret_cast<float *>(smem + Cfg::kSmemXch);  // [2][128] row-max / row-sum exchange
    const uint32_t tail_cols = C - (total_blocks - 1) * kBlockN;   // valid columns in the last key block

    for (uint32_t item = blockIdx.x, it = 0; item < num_items; item += gridDim.x, ++it) {
      const uint32_t split = item % num_splits, tile_item = item / num_splits;  // tile_item = (head, tile)
      // output slot; the scratch fallback lays partials out as [split][head]
      const uint32_t head = tile_item / tiles_per_head + (kFused ? 0u : split * batch);
      const uint32_t q_row0 = (tile_item % tiles_per_head) * kTileM;
      const uint32_t key_block0 = split * num_blocks;
      const uint32_t g0 = it * num_blocks;  // key blocks this CTA has processed before this item (barrier phases)
      float m = -FLT_MAX;  // running max, log2 domain (AttentionKernel+Caching.swift:310); identical in both threads of a row
      float l = 0.f;       // running sum over this thread's columns
      MFA_TRACE(4 + h, it, 0);

      for (uint32_t i = 0; i < num_blocks; ++i) {
        const uint32_t g = g0 + i, bf = g % kSBuffers, ph = (g / kSBuffers) & 1;
        const uint32_t tS = tTile + bf * kBlockN;
        mbar_wait(&s_full[bf], ph);
        tc_fence_after();
        MFA_TRACE(trace_role, i, 0);

        float s[kCols];
        tmem_ld64(tS + h * kCols, *reinterpret_cast<uint32_t(*)[kCols]>(&s[0]));
        tc_wait_ld();
        MFA_TRACE(trace_role, i, 1);

        // edge mask (maskAttentionMatrixEdge, AttentionKernel+Softmax.swift:228-260)
        if (key_block0 + i == total_blocks - 1 && tail_cols < kBlockN) {
#pragma unroll
          for (uint32_t c = 0; c < kCols; ++c)
            if (h * kCols + c >= tail_cols) s[c] = -INFINITY;
        }

        // The reference tracks the exact running row max every block (onlineReduceMaximum / onlineCorrectO,
        // +Softmax.swift:267-301).  Here the max is refreshed lazily: P is computed against the current (possibly
        // stale) m straight away, and only if some half-row of P sums to more than 2^8 -- i.e. some element could
        // exceed 2^8, or m was never set -- do the two warps of these rows fall back to the exact path (joint row max,
        // wait for every issued O += P V, rescale O and l, recompute).  With m lagging the true max by at most 8
        // (log2), P <= 2^8 keeps full FP32 / 16-bit accuracy, and the 128-element max reduction disappears from the
        // common path; the result is mathematically identical.
        uint32_t packed[kCols / 2];
        float half_sum;
        if (i == 0) {
          half_sum = INFINITY;  // m is not set yet: the item's first block goes straight to the exact path
        } else {
          float2 sum2 = make_float2(0.f, 0.f);
          const float2 scale2 = make_float2(scale_log2, scale_log2), negm2 = make_float2(-m, -m);
#pragma unroll
          for (uint32_t k = 0; k < kCols / 2; ++k) {
            const float2 x = ffma2(make_float2(s[2 * k], s[2 * k + 1]), scale2, negm2);
            float2 pr;
            if (kPoly > 0 && (k & 3) < kPoly) {
              pr = exp2_poly2(x);
            } else {
              pr.x = ex2_approx(x.x);
              pr.y = ex2_approx(x.y);
            }
            packed[k] = kBF16 ? pack_bf16x2(pr.x, pr.y) : pack_f16x2(pr.x, pr.y);
            if (kSumRoundedP) pr = kBF16 ? unpack_bf16x2(packed[k]) : unpack_f16x2(packed[k]);
            sum2 = fadd2(sum2, pr);
          }
          half_sum = sum2.x + sum2.y;
        }
        // (the barrier also orders both warps' S loads before either overwrites the buffer with P: warpgroup 1's P
        // columns [32, 64) lie inside warpgroup 0's S columns [0, 64))
        if (bar_red_or(pair_bar, 64, !(half_sum <= kLazySumLimit))) {  // also catches inf / NaN
          // ---- exact path (rare) ----
          float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
          for (uint32_t c = 4; c < kCols; c += 4) {
            mx0 = fmaxf(mx0, s[c]);
            mx1 = fmaxf(mx1, s[c + 1]);
            mx2 = fmaxf(mx2, s[c + 2]);
            mx3 = fmaxf(mx3, s[c + 3]);
          }
          const float m_loc = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
          xch[h * kTileM + row_in_tile] = m_loc;
          bar_sync(pair_bar, 64);
          const float m_new = fmaxf(m, fmaxf(m_loc, xch[(1 - h) * kTileM + row_in_tile]));
          if (i > 0) {
            mbar_wait(o_full, (g - 1) & 1);  // O += P V of the previous block has landed (this block's is not issued yet)
            tc_fence_after();
            const float correction = ex2_approx(m - m_new);
#pragma unroll
            for (uint32_t c = 0; c < DPAD / 2; c += 32) {  // each warpgroup rescales its half of the O columns
              uint32_t o[32];
              tmem_ld32(tO + h * (DPAD / 2) + c, o);
              tc_wait_ld();
#pragma unroll
              for (uint32_t k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * correction);
              tmem_st32(tO + h * (DPAD / 2) + c, o);
            }
            l *= correction;
          }
          m = m_new;
          float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
          for (uint32_t k = 0; k < kCols / 2; ++k) {
            const float p0 = ex2_approx(fmaf(s[2 * k], scale_log2, -m));
            const float p1 = ex2_approx(fmaf(s[2 * k + 1], scale_log2, -m));
            packed[k] = kBF16 ? pack_bf16x2(p0, p1) : pack_f16x2(p0, p1);
            if (kSumRoundedP) {
              const float2 q = kBF16 ? unpack_bf16x2(packed[k]) : unpack_f16x2(packed[k]);
              sum0 += q.x;
              sum1 += q.y;
            } else {
              sum0 += p0;
              sum1 += p1;
            }
          }
          half_sum = sum0 + sum1;
          bar_sync(pair_bar, 64);  // the exchange slots may be rewritten in a later block only after both have read them
        }
        l += half_sum;
        MFA_TRACE(trace_role, i, 2);
        // P (16-bit) over S: keys [64 h, 64 h + 64) -> columns [32 h, 32 h + 32); each warpgroup releases its half on
        // its own barrier, so the MMA warp starts O += P V on whichever 64 keys are ready
        tmem_st32(tS + h * (kCols / 2), packed);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&p_full[2 * bf + h]);
        MFA_TRACE(trace_role, i, 3);
      }

      // ---------------- epilogue: O / l -> global (FP32), L = m + log2(l) ----------------
      // (fused split-KV: the RAW accumulator -> this split's slot of the workspace instead, then the merge below)
      MFA_TRACE(4 + h, it, 1);
      xch[h * kTileM + row_in_tile] = l;
      bar_sync(pair_bar, 64);
      l += xch[(1 - h) * kTileM + row_in_tile];
      mbar_wait(o_final, it & 1);
      tc_fence_after();
      MFA_TRACE(4 + h, it, 2);
      const uint32_t row = q_row0 + row_in_tile;
      const float out_scale = kFused ? 1.0f : 1.0f / l;
      {
        // Warpgroup h stores columns [h DPAD/2, (h+1) DPAD/2).  TMEM hands every thread one row; storing rows straight
        // from registers would touch 32 different cache lines per warp store, so each warp transposes 32 x 32 chunks
        // through a private XOR-swizzled scratch tile (128-bit accesses, conflict-free both ways) and writes four full
        // 128 B lines per store instruction.
        const uint32_t scratch = smem_u32(smem + Cfg::kSmemScratch) + warp * 4096;
        const uint32_t warp_row0 = q_row0 + quarter * 32;
        // rows the stores may touch: the problem's R rows, or all 128 slots of the tile in the workspace
        const uint32_t row_limit = kFused ? q_row0 + kTileM : R;
        float *o_base = kFused ? part_O + ((static_cast<size_t>(tile_item) * num_splits + split) * kTileM + quarter * 32) * D
                               : O + (static_cast<size_t>(head) * R + warp_row0) * D;
        const uint32_t sub_row = lane >> 3, quad = lane & 7;  // transposed view: 4 rows x 8 float4 per warp access
#pragma unroll
        for (uint32_t cc = 0; cc < DPAD / 2; cc += 32) {
          const uint32_t c = h * (DPAD / 2) + cc;
          uint32_t o[32];
          tmem_ld32(tO + c, o);
          tc_wait_ld();
#pragma unroll
          for (uint32_t j = 0; j < 8; ++j)
            sts_f32x4(scratch + (lane * 8 + (j ^ (lane & 7))) * 16,
                      make_float4(__uint_as_float(o[4 * j]) * out_scale, __uint_as_float(o[4 * j + 1]) * out_scale,
                                  __uint_as_float(o[4 * j + 2]) * out_scale, __uint_as_float(o[4 * j + 3]) * out_scale));
          __syncwarp();
          // the whole transposed chunk goes into distinct registers BEFORE the first store: a store keeps its source
          // registers busy until the data has left the SM, so reusing a handful of registers would serialise the
          // stores on memory latency
          float4 v[8];
#pragma unroll
          for (uint32_t k = 0; k < 8; ++k) {
            const uint32_t r = 4 * k + sub_row;
            v[k] = lds_f32x4(scratch + (r * 8 + (quad ^ (r & 7))) * 16);
          }
          if (c + 4 * quad < D) {  // D % 8 == 0: a float4 is either fully inside or fully outside
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) {
              const uint32_t r = 4 * k + sub_row;
              if (warp_row0 + r < row_limit) *reinterpret_cast<float4 *>(o_base + static_cast<size_t>(r) * D + c + 4 * quad) = v[k];
            }
          }
          __syncwarp();
        }
      }
      // O is out of TMEM: the next item's first O = P V (accumulate off) may overwrite it
      tc_fence_before();
      mbar_arrive(o_free);
      MFA_TRACE(4 + h, it, 3);

      if constexpr (!kFused) {
        if (h == 0 && row < R && L != nullptr) {
          const float lse2 = m + log2f(l);  // AttentionKernel+Caching.swift:373-377
          const size_t idx = static_cast<size_t>(head) * R + row;
          if (l_is_fp16)
            reinterpret_cast<__half *>(L)[idx] = __float2half_rn(lse2);
          else
            reinterpret_cast<float *>(L)[idx] = lse2;
        }
      } else {
        // ---------------- fused split-KV: publish the partial, wait for the siblings, merge a slice of the rows ------
        const size_t slot0 = static_cast<size_t>(tile_item) * num_splits * kTileM;  // first row slot of this tile
        if (h == 0) part_ml[slot0 + static_cast<size_t>(split) * kTileM + row_in_tile] = make_float2(m, l);
        __threadfence();  // the partial is visible device-wide before the arrival is
        bar_sync(1, kSoftmaxThreads);
        uint32_t *arrived = counters + 2 * tile_item, *merged = arrived + 1;
        if (threadIdx.x == 0) {
          atomicAdd(arrived, 1u);
          // bounded spin (a protocol bug must trap, not hang): the siblings are co-resident, so this is short
          const long long start = clock64();
          uint32_t seen;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(arrived) : "memory");
            if (seen < num_splits && clock64() - start > MFA_MBAR_TIMEOUT_CYCLES) {
              printf("mfa_b200: split-KV arrival timeout block %d (%u of %u)\n", blockIdx.x, seen, num_splits);
              __trap();
            }
          } while (seen < num_splits);
        }
        bar_sync(1, kSoftmaxThreads);
        MFA_TRACE(4 + h, it, 4);
        // rows [split * rows_per_rank, ...) of the tile: thread <-> (row, 16 B column slot), DPAD / 4 consecutive threads
        // per row, so the global stores are full lines; every load of a thread is issued before the first use
        constexpr uint32_t kQuadsPerRow = DPAD / 4;
        const uint32_t rows_per_rank = (kTileM + num_splits - 1) / num_splits;
        const uint32_t mq = threadIdx.x % kQuadsPerRow;
        for (uint32_t r = threadIdx.x / kQuadsPerRow; r < rows_per_rank; r += kSoftmaxThreads / kQuadsPerRow) {
          const uint32_t rt = split * rows_per_rank + r;
          if (rt >= kTileM) break;
          float2 mls[16];
          float4 vs[16];
#pragma unroll
          for (uint32_t sp = 0; sp < 16; ++sp) {
            mls[sp] = make_float2(-FLT_MAX, 0.f);
            vs[sp] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (sp < num_splits) {
              const size_t slot = slot0 + static_cast<size_t>(sp) * kTileM + rt;
              mls[sp] = __ldcg(part_ml + slot);
              if (4 * mq < D) vs[sp] = __ldcg(reinterpret_cast<const float4 *>(part_O + slot * D) + mq);
            }
          }
          float m_all = -FLT_MAX;
#pragma unroll
          for (uint32_t sp = 0; sp < 16; ++sp) m_all = fmaxf(m_all, mls[sp].x);
          float denom = 0.f;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (uint32_t sp = 0; sp < 16; ++sp) {
            const float w = exp2f(mls[sp].x - m_all);  // unused slots: l = 0 and O = 0, so their weight is irrelevant
            denom = fmaf(w, mls[sp].y, denom);
            acc.x = fmaf(w, vs[sp].x, acc.x);
            acc.y = fmaf(w, vs[sp].y, acc.y);
            acc.z = fmaf(w, vs[sp].z, acc.z);
            acc.w = fmaf(w, vs[sp].w, acc.w);
          }
          const float inv = 1.0f / denom;
          const uint32_t out_row = q_row0 + rt;
          if (out_row < R) {
            if (4 * mq < D)
              *reinterpret_cast<float4 *>(O + (static_cast<size_t>(head) * R + out_row) * D + 4 * mq) =
                  make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
            if (mq == 0 && L != nullptr) {
              const float lse2 = m_all + log2f(denom);  // AttentionKernel+Caching.swift:373-377
              const size_t idx = static_cast<size_t>(head) * R + out_row;
              if (l_is_fp16)
                reinterpret_cast<__half *>(L)[idx] = __float2half_rn(lse2);
              else
                reinterpret_cast<float *>(L)[idx] = lse2;
            }
          }
        }
        // the last CTA to finish reading returns both counters to zero for the next launch on this stream
        bar_sync(1, kSoftmaxThreads);
        if (threadIdx.x == 0 && atomicAdd(merged, 1u) == num_splits - 1) {
          *arrived = 0;
          *merged = 0;
        }
        MFA_TRACE(4 + h, it, 5);
      }
    }  // work items
  } else {
    setmaxnreg_dec<kOtherRegs>();
    // The producer warps run their control flow warp-wide and hand exactly one elected lane to the
    // TMA / tcgen05 instructions: operands stay in uniform registers and the issue loops are branch-free.
    if (warp == 9) {
      // ===================================================================================
      // TMA producer: Q of every item (two buffers: the next item's Q lands during the current item) and the K ring
      // ===================================================================================
      for (uint32_t item = blockIdx.x, it = 0; item < num_items; item += gridDim.x, ++it) {
        const uint32_t tile_item = item / num_splits;
        const uint32_t head = tile_item / tiles_per_head;
        const uint32_t q_row0 = (tile_item % tiles_per_head) * kTileM;
        const uint32_t key_block0 = (item % num_splits) * num_blocks;
        const uint32_t g0 = it * num_blocks, qb = it & 1;
        mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);  // item it - 2's S MMAs are done with this Q buffer
        if (elect_one()) {
          mbar_arrive_expect_tx(&q_full[qb], Cfg::kTileBytes);
#pragma unroll
          for (uint32_t ds = 0; ds < Cfg::kSubTiles; ++ds)
            tma_load_3d(smem + Cfg::kSmemQ + qb * Cfg::kTileBytes + ds * kSubTileBytes, &mapQ, &q_full[qb], ds * 64, q_row0, head);
        }
        for (uint32_t i = 0; i < num_blocks; ++i) {
          const uint32_t stage = (g0 + i) % Cfg::kStages, phase = ((g0 + i) / Cfg::kStages) & 1;
          mbar_wait(&k_empty[stage], phase ^ 1);
          MFA_TRACE(3, i, 0);
          if (elect_one()) {
            mbar_arrive_expect_tx(&k_full[stage], Cfg::kTileBytes);
#pragma unroll
            for (uint32_t ds = 0; ds < Cfg::kSubTiles; ++ds)
              tma_load_3d(smem + Cfg::kSmemK + stage * Cfg::kTileBytes + ds * kSubTileBytes, &mapK, &k_full[stage],
                          ds * 64, (key_block0 + i) * kBlockN, head);
          }
        }
      }
    } else if (warp == 10) {
      // ===================================================================================
      // TMA producer for the V ring (a separate warp: K loads must not queue behind the wait for a V stage)
      // ===================================================================================
      for (uint32_t item = blockIdx.x, it = 0; item < num_items; item += gridDim.x, ++it) {
        const uint32_t head = (item / num_splits) / tiles_per_head;
        const uint32_t key_block0 = (item % num_splits) * num_blocks;
        const uint32_t g0 = it * num_blocks;
        for (uint32_t i = 0; i < num_blocks; ++i) {
          const uint32_t stage = (g0 + i) % Cfg::kStages, phase = ((g0 + i) / Cfg::kStages) & 1;
          mbar_wait(&v_empty[stage], phase ^ 1);
          MFA_TRACE(3, i, 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&v_full[stage], Cfg::kTileBytes);
#pragma unroll
            for (uint32_t ds = 0; ds < Cfg::kSubTiles; ++ds)
              tma_load_3d(smem + Cfg::kSmemV + stage * Cfg::kTileBytes + ds * kSubTileBytes, &mapV, &v_full[stage],
                          ds * 64, (key_block0 + i) * kBlockN, head);
          }
        }
      }
    } else if (warp == 8) {
      // ===================================================================================
      // MMA issuer
      // ===================================================================================
      constexpr uint32_t kFormat = kBF16 ? 1u : 0u;
      // S[128 x 128] = Q[128 x D] . K[128 x D]^T : A and B both K-major
      constexpr uint32_t idescS = make_idesc_f16(kTileM, kBlockN, kFormat, 0, 0);
      // O[128 x DPAD] += P[128 x 128] . V[128 x DPAD] : A from TMEM, B (= V, [key][d]) is MN-major
      constexpr uint32_t idescO = make_idesc_f16(kTileM, DPAD, kFormat, 0, 1);
      // Descriptors differ only in the 14-bit start-address field; build each once and add (bytes >> 4).
      const uint64_t descQ = make_smem_desc_sw128(smem_u32(smem + Cfg::kSmemQ), 16, 1024);
      const uint64_t descK = make_smem_desc_sw128(smem_u32(smem + Cfg::kSmemK), 16, 1024);
      const uint64_t descV = make_smem_desc_sw128(smem_u32(smem + Cfg::kSmemV), kSubTileBytes, 1024);

      // every tcgen05.mma / commit below is issued by the one elected lane
      auto issue_S = [&](uint32_t bf, uint32_t qb, uint32_t stage) {
        const uint32_t d_tmem = tmem_base + bf * kBlockN;
        const uint64_t a0 = descQ + ((qb * Cfg::kTileBytes) >> 4);
        const uint64_t b0 = descK + ((stage * Cfg::kTileBytes) >> 4);
#pragma unroll
        for (uint32_t k = 0; k < DPAD / 16; ++k) {
          // 16 elements = 32 B inside the 128 B swizzle row; 4 k-steps per 64-element sub-tile
          const uint32_t off = ((k >> 2) * kSubTileBytes + (k & 3) * 32) >> 4;
          umma_ss(d_tmem, a0 + off, b0 + off, idescS, k > 0);
        }
      };
      // half 0: keys 0-63 of the block (P columns 0-31 of the buffer), half 1: keys 64-127
      auto issue_PV = [&](uint32_t bf, uint32_t half, uint32_t stage, uint32_t accumulate) {
        const uint32_t d_tmem = tmem_base + Cfg::kTmemO;
        const uint32_t a_tmem = tmem_base + bf * kBlockN;
        const uint64_t b0 = descV + ((stage * Cfg::kTileBytes) >> 4);
#pragma unroll
        for (uint32_t kk = 0; kk < kBlockN / 32; ++kk) {
          const uint32_t k = half * (kBlockN / 32) + kk;
          // 16 keys = two 8-row groups of 1024 B; 64-wide column blocks are kSubTileBytes apart (LBO)
          umma_ts(d_tmem, a_tmem + k * 8, b0 + ((k * 2048) >> 4), idescO, k > 0 ? 1u : accumulate);
        }
      };

      for (uint32_t item = blockIdx.x, it = 0; item < num_items; item += gridDim.x, ++it) {
        const uint32_t g0 = it * num_blocks, qb = it & 1;
        MFA_TRACE(6, it, 0);
        mbar_wait(&q_full[qb], (it >> 1) & 1);
        // prologue: the item's first kSBuffers S tiles.  They overlap the softmax warps' epilogue of the previous item
        // (the S / P buffers are free once the previous item's last O += P V has been issued: the pipe runs in order).
        for (uint32_t i = 0; i < kSBuffers && i < num_blocks; ++i) {
          const uint32_t g = g0 + i, stage = g % Cfg::kStages;
          mbar_wait(&k_full[stage], (g / Cfg::kStages) & 1);
          tc_fence_after();
          if (elect_one()) {
            issue_S(g % kSBuffers, qb, stage);
            umma_commit(&s_full[g % kSBuffers]);
            umma_commit(&k_empty[stage]);
            if (i + 1 == num_blocks) umma_commit(&q_empty[qb]);
          }
          __syncwarp();
        }
        MFA_TRACE(6, it, 1);
        for (uint32_t i = 0; i < num_blocks; ++i) {
          const uint32_t g = g0 + i, bf = g % kSBuffers, ph = (g / kSBuffers) & 1;
          const uint32_t stage = g % Cfg::kStages;
          mbar_wait(&v_full[stage], (g / Cfg::kStages) & 1);
          // the previous item's epilogue must have read O out before accumulate-off overwrites it
          if (i == 0 && it > 0) mbar_wait(o_free, (it - 1) & 1);
          MFA_TRACE(2, i, 0);
          mbar_wait(&p_full[2 * bf], ph);
          tc_fence_after();
          MFA_TRACE(2, i, 1);
          if (elect_one()) issue_PV(bf, 0, stage, i > 0 ? 1u : 0u);
          __syncwarp();
          mbar_wait(&p_full[2 * bf + 1], ph);
          tc_fence_after();
          MFA_TRACE(2, i, 2);
          if (elect_one()) {
            issue_PV(bf, 1, stage, 1u);
            umma_commit(o_full);
            umma_commit(&v_empty[stage]);
            if (i + 1 == num_blocks) umma_commit(o_final);
          }
          __syncwarp();
          // S(i + kSBuffers) overwrites P(i) only after O += P V (i): the tensor pipe runs in order
          if (i + kSBuffers < num_blocks) {
            const uint32_t gn = g + kSBuffers, nstage = gn % Cfg::kStages;
            mbar_wait(&k_full[nstage], (gn / Cfg::kStages) & 1);
            tc_fence_after();
            if (elect_one()) {
              issue_S(bf, qb, nstage);
              umma_commit(&s_full[bf]);
              umma_commit(&k_empty[nstage]);
              if (i + kSBuffers + 1 == num_blocks) umma_commit(&q_empty[qb]);  // that was the item's last read of Q
            }
            __syncwarp();
          }
          MFA_TRACE(2, i, 3);
        }
        MFA_TRACE(6, it, 2);
      }  // work items
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

// Merges the num_splits partial results of split-KV (scratch form): L = log2 sum_s 2^L_s,  O = sum_s 2^(L_s - L) O_s.
// One thread per (row, 4 columns); partials are [split][head][row][D] FP32 and [split][head][row] FP32.  Every load of
// a thread is issued before the first use (the partials sit in L2; a dependent chain of num_splits round trips was
// 3x slower).  Launched with programmatic stream serialisation: the grid is set up while the attention kernel drains
// and griddepcontrol.wait holds it until that kernel's writes are visible.
template <uint32_t kMaxSplits>
__global__ void __launch_bounds__(128)
    combine_splits(const float *__restrict__ O_part, const float *__restrict__ L_part, float *__restrict__ O,
                   void *__restrict__ L, uint32_t rows_total, uint32_t D, uint32_t num_splits, int l_is_fp16) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t quads_per_row = D / 4;
  const uint64_t idx = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint64_t row = idx / quads_per_row;
  const uint32_t quad = static_cast<uint32_t>(idx % quads_per_row);
  if (row >= rows_total) return;
  float ls[kMaxSplits];
  float4 v[kMaxSplits];
#pragma unroll
  for (uint32_t s = 0; s < kMaxSplits; ++s) {
    ls[s] = -INFINITY;
    v[s] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s < num_splits) {
      ls[s] = __ldcg(L_part + static_cast<uint64_t>(s) * rows_total + row);
      v[s] = __ldcg(reinterpret_cast<const float4 *>(O_part + (static_cast<uint64_t>(s) * rows_total + row) * D) + quad);
    }
  }
  float lmax = ls[0];
#pragma unroll
  for (uint32_t s = 1; s < kMaxSplits; ++s) lmax = fmaxf(lmax, ls[s]);
  float denom = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (uint32_t s = 0; s < kMaxSplits; ++s) {
    const float w = exp2f(ls[s] - lmax);  // 0 for the unused slots
    denom += w;
    acc.x = fmaf(w, v[s].x, acc.x);
    acc.y = fmaf(w, v[s].y, acc.y);
    acc.z = fmaf(w, v[s].z, acc.z);
    acc.w = fmaf(w, v[s].w, acc.w);
  }
  const float inv = 1.0f / denom;
  *reinterpret_cast<float4 *>(O + row * D + 4 * quad) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  if (quad == 0 && L != nullptr) {
    const float lse = lmax + log2f(denom);
    if (l_is_fp16)
      reinterpret_cast<__half *>(L)[row] = __float2half_rn(lse);
    else
      reinterpret_cast<float *>(L)[row] = lse;
  }
}

// how many key ranges to cut every item into: only when the SMs would otherwise idle, only into equal ranges of at
// least four key blocks (shorter ranges are dominated by the per-item prologue / epilogue)
static uint32_t choose_splits(uint32_t items, uint32_t total_blocks, uint32_t sm_count, uint32_t max_splits = 16) {
  if (items * 2 > sm_count) return 1;
  const uint32_t target = sm_count / items;
  uint32_t best = 1;
  for (uint32_t s = 2; s <= target && s <= max_splits; ++s)
    if (total_blocks % s == 0 && total_blocks / s >= 4) best = s;
  return best;
}

template <uint32_t DPAD, bool kBF16, bool kTrace = false>
cudaError_t launch(const AttentionParams &p, cudaStream_t stream, long long *trace = nullptr) {
  constexpr uint32_t kSB = MFA_FWD2_SBUFFERS;
  constexpr uint32_t kPoly = DPAD <= 64 ? MFA_FWD2_POLY_D64 : MFA_FWD2_POLY_D128;
  using Cfg = Config<DPAD, kSB>;
  auto kernel = attention_forward_tcgen05<DPAD, kBF16, kSB, kPoly, false, kTrace>;
  const int device = current_device();
  cudaError_t e;
  if ((e = ensure_max_dynamic_smem(reinterpret_cast<const void *>(kernel), Cfg::kSmemBytes, device)) != cudaSuccess) return e;

  CUtensorMap mapQ, mapK, mapV;
  if ((e = make_tensor_map_16bit(&mapQ, p.buf[sQ], p.R, p.D, p.batch, kTileM)) != cudaSuccess) return e;
  if ((e = make_tensor_map_16bit(&mapK, p.buf[sK], p.C, p.D, p.batch, kBlockN)) != cudaSuccess) return e;
  if ((e = make_tensor_map_16bit(&mapV, p.buf[sV], p.C, p.D, p.batch, kBlockN)) != cudaSuccess) return e;

  const uint32_t tiles_per_head = (p.R + kTileM - 1) / kTileM;
  const uint32_t num_items = tiles_per_head * p.batch;
  const uint32_t sm_count = device_sm_count(device);
  const uint32_t total_blocks = (p.C + kBlockN - 1) / kBlockN;
  const int l_is_fp16 = p.prec[sL] == FP16 ? 1 : 0;

  const uint32_t splits = choose_splits(num_items, total_blocks, sm_count);
  if (splits == 1) {
    const uint32_t grid = num_items < sm_count ? num_items : sm_count;
    kernel<<<grid, kThreads, Cfg::kSmemBytes, stream>>>(mapQ, mapK, mapV, static_cast<float *>(p.buf[sO]), p.buf[sL],
                                                        p.R, p.C, p.D, p.scale_log2, l_is_fp16, num_items,
                                                        tiles_per_head, 1u, p.batch, nullptr, nullptr, nullptr, trace);
    return cudaGetLastError();
  }

  // ---- split-KV: partials live in the library's per-(device, stream) workspace -----------------------------------
  const uint32_t split_items = num_items * splits;  // <= sm_count by construction of choose_splits
  if (g_forward_fused_enabled && split_items <= sm_count && 2 * num_items * sizeof(uint32_t) <= kWorkspaceCounterBytes) {
    // fused form: one cooperative launch (every CTA resident, one item each); [counters | O partials | (m, l)]
    auto fused = attention_forward_tcgen05<DPAD, kBF16, kSB, kPoly, true, kTrace>;
    if ((e = ensure_max_dynamic_smem(reinterpret_cast<const void *>(fused), Cfg::kSmemBytes, device)) != cudaSuccess) return e;
    const size_t slots = static_cast<size_t>(split_items) * kTileM;
    const size_t o_bytes = slots * p.D * sizeof(float), ml_bytes = slots * sizeof(float2);
    void *ws = nullptr;
    if ((e = workspace_for(device, stream, o_bytes + ml_bytes, &ws)) != cudaSuccess) return e;
    uint32_t *counters = static_cast<uint32_t *>(ws);
    float *part_O = reinterpret_cast<float *>(static_cast<char *>(ws) + kWorkspaceCounterBytes);
    float2 *part_ml = reinterpret_cast<float2 *>(reinterpret_cast<char *>(part_O) + o_bytes);
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeCooperative;
    attr.val.cooperative = 1;
    cudaLaunchConfig_t config = {};
    config.gridDim = dim3(split_items, 1, 1);
    config.blockDim = dim3(kThreads, 1, 1);
    config.dynamicSmemBytes = Cfg::kSmemBytes;
    config.stream = stream;
    config.attrs = &attr;
    config.numAttrs = 1;
    e = cudaLaunchKernelEx(&config, fused, mapQ, mapK, mapV, static_cast<float *>(p.buf[sO]), p.buf[sL], p.R, p.C, p.D,
                           p.scale_log2, l_is_fp16, split_items, tiles_per_head, splits, p.batch, part_O, part_ml,
                           counters, trace);
    if (e == cudaSuccess) return cudaGetLastError();
    cudaGetLastError();  // cooperative launch refused (e.g. a partitioned GPU): fall through to the two-launch form
  }

  // scratch form: normalised partial O / L per split ([split][head][row]), then the combine kernel
  const uint64_t rows_total = static_cast<uint64_t>(p.batch) * p.R;
  const size_t o_bytes = static_cast<size_t>(splits) * rows_total * p.D * sizeof(float);
  const size_t l_bytes = static_cast<size_t>(splits) * rows_total * sizeof(float);
  void *ws = nullptr;
  if ((e = workspace_for(device, stream, o_bytes + l_bytes, &ws)) != cudaSuccess) return e;
  float *scratch = reinterpret_cast<float *>(static_cast<char *>(ws) + kWorkspaceCounterBytes);
  float *L_part = scratch + static_cast<size_t>(splits) * rows_total * p.D;
  const uint32_t grid = split_items < sm_count ? split_items : sm_count;
  kernel<<<grid, kThreads, Cfg::kSmemBytes, stream>>>(mapQ, mapK, mapV, scratch, L_part, p.R, p.C, p.D, p.scale_log2,
                                                      0, split_items, tiles_per_head, splits, p.batch, nullptr, nullptr,
                                                      nullptr, trace);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  const uint64_t threads = rows_total * (p.D / 4);
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t config = {};
  config.gridDim = dim3(static_cast<uint32_t>((threads + 127) / 128), 1, 1);
  config.blockDim = dim3(128, 1, 1);
  config.stream = stream;
  config.attrs = &attr;
  config.numAttrs = 1;
  auto combine = splits <= 4 ? combine_splits<4> : (splits <= 8 ? combine_splits<8> : combine_splits<16>);
  e = cudaLaunchKernelEx(&config, combine, static_cast<const float *>(scratch), static_cast<const float *>(L_part),
                         static_cast<float *>(p.buf[sO]), p.buf[sL], static_cast<uint32_t>(rows_total), p.D, splits,
                         l_is_fp16);
  return e == cudaSuccess ? cudaGetLastError() : e;
}

}  // namespace fwd2

uint32_t tcgen05_forward_max_head() { return 256; }

// Transposed operands are served by the layout-generic kernel (tcgen05_forward_d256.cu) when TMA can address them: a
// transposed operand's row pitch is its sequence length, which must then be a multiple of 8 elements (16 bytes).
bool tcgen05_forward_transposes_ok(uint32_t R, uint32_t C, bool tQ, bool tK, bool tV) {
  return (!tQ || R % 8 == 0) && (!tK || C % 8 == 0) && (!tV || C % 8 == 0);
}

bool tcgen05_forward_supported(const AttentionParams &p) {
  return (p.prec[sQ] == FP16 || p.prec[sQ] == BF16) && p.prec[sK] == p.prec[sQ] && p.prec[sV] == p.prec[sQ] &&
         p.prec[sO] == FP32 && p.D % 8 == 0 && p.D <= tcgen05_forward_max_head() &&
         tcgen05_forward_transposes_ok(p.R, p.C, p.transposed[sQ], p.transposed[sK], p.transposed[sV]);
}

cudaError_t launch_tcgen05_forward(const AttentionParams &p, cudaStream_t stream) {
  if (!tcgen05_forward_supported(p)) {
    set_launch_detail("descriptor is outside the tcgen05 forward kernel's domain");
    return cudaErrorInvalidValue;
  }
  if (p.transposed[sQ] || p.transposed[sK] || p.transposed[sV] || p.transposed[sO])
    return launch_tcgen05_forward_generic(p, stream);  // tcgen05_forward_d256.cu
  const bool bf16 = p.prec[sQ] == BF16;
  if (p.D > 128) return launch_tcgen05_forward_d256(p, stream);  // tcgen05_forward_d256.cu
  if (p.D <= 64) return bf16 ? fwd2::launch<64, true>(p, stream) : fwd2::launch<64, false>(p, stream);
  return bf16 ? fwd2::launch<128, true>(p, stream) : fwd2::launch<128, false>(p, stream);
}

// Debug entry (not in include/mfa_b200.h): the D=128 bf16 forward with pipeline timestamps of CTA 0
// (8 roles x 64 iterations x 8 slots of clock64()).  Used by scripts/trace_forward.py.
cudaError_t launch_tcgen05_forward_trace(const AttentionParams &p, cudaStream_t stream, long long *trace) {
  return fwd2::launch<128, true, true>(p, stream, trace);
}

// 1 launch, or 2 (attention + combine) when the scratch form of split-KV engages for this problem size
uint32_t tcgen05_forward_launch_count(uint32_t R, uint32_t C, uint32_t D, uint32_t batch) {
  if (D > 128) return 1;
  const uint32_t sm_count = device_sm_count(current_device());
  const uint32_t tiles = (R + fwd2::kTileM - 1) / fwd2::kTileM;
  const uint32_t blocks = (C + fwd2::kBlockN - 1) / fwd2::kBlockN;
  if (fwd2::choose_splits(tiles * batch, blocks, sm_count) == 1) return 1;
  return g_forward_fused_enabled ? 1 : 2;  // fused split-KV merges inside the attention kernel
}

void tcgen05_forward_geometry(uint32_t D, uint32_t *threads, uint32_t *smem_bytes, uint32_t *par, uint32_t *trav,
                              uint32_t *head) {
  if (D > 128) {
    tcgen05_forward_d256_geometry(threads, smem_bytes, par, trav);
    *head = 256 < (D + 7) / 8 * 8 ? 256 : (D + 7) / 8 * 8;
    return;
  }
  *threads = fwd2::kThreads;
  *smem_bytes = D <= 64 ? fwd2::Config<64, MFA_FWD2_SBUFFERS>::kSmemBytes : fwd2::Config<128, MFA_FWD2_SBUFFERS>::kSmemBytes;
  *par = fwd2::kTileM;
  *trav = fwd2::kBlockN;
  *head = D <= 64 ? 64 : 128;
  const uint32_t padded = (D + 7) / 8 * 8;
  if (*head > padded) *head = padded;
}

}  // namespace mfa
