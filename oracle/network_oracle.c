/*
 * network_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, FP32 scalar, same loop order) of the reference's only CPU
 * attention implementation, `struct Network`
 * (/root/reference/Tests/FlashAttentionTests/Utilities/Network.swift:70-403), which the
 * reference's own tests use as the correctness oracle for the forward / dQ / dK-dV kernels.
 *
 * Who may use this file: tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline /
 * --impl reference legs -- only ever as the checker or the timed CPU baseline.  The product
 * path (metal-flash-attention_b200/) never links, imports or calls anything in oracle/.
 *
 * PARITY PIN STATUS -- "parity unpinned" in the strict sense: the reference ships NO golden
 * vectors, known-answer tests or fixtures for the attention path (its tests compare the Metal
 * kernels against Network on *unseeded* random inputs, SquareAttentionTest.swift:214-555), and
 * the reference itself cannot run here (Swift + Metal + Apple GPU; no swiftc in this image).
 * What pins this restatement instead (see tests/test_oracle.py, DESIGN.md "Oracle"):
 *   - an independent float64 numpy formulation (oracle/oracle_np.py) agrees to FP32 round-off;
 *   - central finite differences of the reference's loss  Phi = sum dO*O  (Network.swift:314-326,
 *     the method of Documentation/Archive/FiniteDifferencingTest.swift:85-134) agree with
 *     derivativeQ/K/V;
 *   - softmax identities (rows of P sum to 1, L == logsumexp, D == rowsum(dO*O)).
 *
 * The *_omp entry points use the same per-row arithmetic but distribute rows over OpenMP
 * threads (used to generate goldens at large N and as the "all host cores" CPU baseline).
 *
 * Build: see oracle/Makefile  (gcc -O3 -march=x86-64-v3 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Seeded inputs.  Network.init (Network.swift:80-113) draws one Box-Muller pair per element:
 * (Q, dO) share a pair, (K, V) share a pair (boxMullerTransform, :115-129).  The reference's
 * uniform source is unseeded (SIMD2<Float>.random); we substitute splitmix64 so runs repeat.
 * ------------------------------------------------------------------------------------------ */
static uint64_t splitmix64(uint64_t *state) {
  uint64_t z = (*state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

/* uniform in (0,1): 24 random bits, never exactly 0 so log() stays finite */
static float uniform01(uint64_t *state) {
  uint32_t bits = (uint32_t)(splitmix64(state) >> 40); /* 24 bits */
  return ((float)bits + 0.5f) * (1.0f / 16777216.0f);
}

/* Network.swift:115-129 */
static void box_muller(uint64_t *state, float out[2]) {
  float u0 = uniform01(state);
  float u1 = uniform01(state);
  float logPart = logf(u0);
  float magnitudePart = sqrtf(-2.0f * logPart);
  float anglePart = 2.0f * 3.14159265358979323846f * u1;
  out[0] = magnitudePart * cosf(anglePart);
  out[1] = magnitudePart * sinf(anglePart);
}

/* Network.swift:80-113 -- Q,dO are R x D; K,V are C x D; row-major */
ORACLE_API void oracle_network_init(int R, int C, int D, uint64_t seed,
                                    float *Q, float *K, float *V, float *dO) {
  uint64_t state = seed * 0x2545F4914F6CDD1Dull + 0x1234567ull;
  float pair[2];
  for (int rowID = 0; rowID < R; ++rowID) {
    for (int d = 0; d < D; ++d) {
      size_t address = (size_t)rowID * D + d;
      box_muller(&state, pair);
      Q[address] = pair[0];
      dO[address] = pair[1];
    }
  }
  for (int columnID = 0; columnID < C; ++columnID) {
    for (int d = 0; d < D; ++d) {
      size_t address = (size_t)columnID * D + d;
      box_muller(&state, pair);
      K[address] = pair[0];
      V[address] = pair[1];
    }
  }
}

typedef struct {
  int R, C, D;
  const float *Q, *K, *V, *dO;
} Network;

/* Network.swift:134-149  createMatrixSRow: S[row, :] = Q[row, :] . K^T (unscaled) */
static void createMatrixSRow(const Network *n, int rowID, float *output) {
  const int D = n->D;
  for (int columnID = 0; columnID < n->C; ++columnID) {
    float dotProduct = 0.0f;
    for (int d = 0; d < D; ++d) {
      dotProduct += n->Q[(size_t)rowID * D + d] * n->K[(size_t)columnID * D + d];
    }
    output[columnID] = dotProduct;
  }
}

/* Network.swift:151-179  createMatrixPRow: P = exp(s*S - lse), lse = max + log(sum exp) */
static void createMatrixPRow(const Network *n, int rowID, float *output) {
  createMatrixSRow(n, rowID, output);
  const float scaleFactor = 1.0f / sqrtf((float)n->D);

  float maximum = -FLT_MAX;
  for (int columnID = 0; columnID < n->C; ++columnID) {
    float value = scaleFactor * output[columnID];
    maximum = fmaxf(maximum, value);
  }
  float sum = 0.0f;
  for (int columnID = 0; columnID < n->C; ++columnID) {
    float value = scaleFactor * output[columnID];
    sum += expf(value - maximum);
  }
  float lse = maximum + logf(sum);
  for (int columnID = 0; columnID < n->C; ++columnID) {
    float value = scaleFactor * output[columnID];
    output[columnID] = expf(value - lse);
  }
}

/* Network.swift:181-203  createLTerm: natural-log LSE of the scaled row */
static float createLTerm(const Network *n, int rowID, float *scratchS) {
  createMatrixSRow(n, rowID, scratchS);
  const float scaleFactor = 1.0f / sqrtf((float)n->D);
  float maximum = -FLT_MAX;
  for (int columnID = 0; columnID < n->C; ++columnID) {
    maximum = fmaxf(maximum, scaleFactor * scratchS[columnID]);
  }
  float sum = 0.0f;
  for (int columnID = 0; columnID < n->C; ++columnID) {
    sum += expf(scaleFactor * scratchS[columnID] - maximum);
  }
  return maximum + logf(sum);
}

/* Network.swift:205-218  createDerivativePRow: dP[row, :] = dO[row, :] . V^T */
static void createDerivativePRow(const Network *n, int rowID, float *output) {
  const int D = n->D;
  for (int columnID = 0; columnID < n->C; ++columnID) {
    float dotProduct = 0.0f;
    for (int d = 0; d < D; ++d) {
      dotProduct += n->dO[(size_t)rowID * D + d] * n->V[(size_t)columnID * D + d];
    }
    output[columnID] = dotProduct;
  }
}

/* shared by createDerivativeSRow / createDTerm / inferenceAttention: O[row,:] = P[row,:] . V
 * (Network.swift:223-234, 262-273, 291-303: d outer, column inner) */
static void rowPV(const Network *n, const float *matrixPRow, float *matrixORow) {
  const int D = n->D;
  for (int d = 0; d < D; ++d) {
    float dotProduct = 0.0f;
    for (int columnID = 0; columnID < n->C; ++columnID) {
      dotProduct += matrixPRow[columnID] * n->V[(size_t)columnID * D + d];
    }
    matrixORow[d] = dotProduct;
  }
}

/* Network.swift:259-281  createDTerm: D[row] = sum_d O[row,d] * dO[row,d]  (unscaled) */
static float createDTerm(const Network *n, int rowID, float *scratchP, float *scratchO) {
  createMatrixPRow(n, rowID, scratchP);
  rowPV(n, scratchP, scratchO);
  float termD = 0.0f;
  for (int d = 0; d < n->D; ++d) {
    termD += scratchO[d] * n->dO[(size_t)rowID * n->D + d];
  }
  return termD;
}

/* Network.swift:220-257  createDerivativeSRow: dS = P * (dP - D) * (1/sqrt(D)) */
static void createDerivativeSRow(const Network *n, int rowID, float *derivativeSRow,
                                 float *scratchP, float *scratchO, float *scratchdP) {
  float termD = createDTerm(n, rowID, scratchP, scratchO);
  createDerivativePRow(n, rowID, scratchdP);
  const float scaleFactor = 1.0f / sqrtf((float)n->D);
  for (int columnID = 0; columnID < n->C; ++columnID) {
    float valueP = scratchP[columnID];
    float valueDerivativeP = scratchdP[columnID];
    float valueS = valueP * (valueDerivativeP - termD);
    valueS *= scaleFactor;
    derivativeSRow[columnID] = valueS;
  }
}

static Network make_network(int R, int C, int D, const float *Q, const float *K,
                            const float *V, const float *dO) {
  Network n;
  n.R = R; n.C = C; n.D = D; n.Q = Q; n.K = K; n.V = V; n.dO = dO;
  return n;
}

/* Network.swift:286-311  inferenceAttention -> O [R x D].  Also returns L (createLTerm, natural
 * log) when L != NULL.  threads <= 1: the reference's single-thread order. */
ORACLE_API void oracle_inference_attention(int R, int C, int D, const float *Q, const float *K,
                                           const float *V, float *O, float *L, int threads) {
  Network n = make_network(R, C, D, Q, K, V, NULL);
#pragma omp parallel num_threads(threads > 1 ? threads : 1) if (threads > 1)
  {
    float *matrixPRow = (float *)malloc(sizeof(float) * (size_t)C);
    float *matrixORow = (float *)malloc(sizeof(float) * (size_t)D);
#pragma omp for schedule(static)
    for (int rowID = 0; rowID < R; ++rowID) {
      createMatrixPRow(&n, rowID, matrixPRow);
      rowPV(&n, matrixPRow, matrixORow);
      for (int d = 0; d < D; ++d) O[(size_t)rowID * D + d] = matrixORow[d];
    }
    if (L) {
#pragma omp for schedule(static)
      for (int rowID = 0; rowID < R; ++rowID) L[rowID] = createLTerm(&n, rowID, matrixPRow);
    }
    free(matrixPRow);
    free(matrixORow);
  }
}

/* (0..<R).map(createDTerm)  -- SquareAttentionTest.swift:397 */
ORACLE_API void oracle_d_terms(int R, int C, int D, const float *Q, const float *K,
                               const float *V, const float *dO, float *Dterm, int threads) {
  Network n = make_network(R, C, D, Q, K, V, dO);
#pragma omp parallel num_threads(threads > 1 ? threads : 1) if (threads > 1)
  {
    float *scratchP = (float *)malloc(sizeof(float) * (size_t)C);
    float *scratchO = (float *)malloc(sizeof(float) * (size_t)D);
#pragma omp for schedule(static)
    for (int rowID = 0; rowID < R; ++rowID) Dterm[rowID] = createDTerm(&n, rowID, scratchP, scratchO);
    free(scratchP);
    free(scratchO);
  }
}

/* Network.swift:314-326  loss: Phi = sum_n sum_d dO[n][d] * O[n][d] (accumulated in double here
 * only because it feeds finite differences in tests; the reference accumulates in Float) */
ORACLE_API double oracle_loss(int R, int C, int D, const float *Q, const float *K, const float *V,
                              const float *dO) {
  float *O = (float *)malloc(sizeof(float) * (size_t)R * D);
  oracle_inference_attention(R, C, D, Q, K, V, O, NULL, 1);
  double output = 0.0;
  for (size_t i = 0; i < (size_t)R * D; ++i) output += (double)dO[i] * (double)O[i];
  free(O);
  return output;
}

/* Network.swift:329-349  derivativeV: dV = P^T dO, row-outer accumulation order.
 * threads > 1: rows are split over threads into private accumulators that are then summed in
 * thread order (summation order differs from the single-thread reference order). */
ORACLE_API void oracle_derivative_v(int R, int C, int D, const float *Q, const float *K,
                                    const float *V, const float *dO, float *dV, int threads) {
  Network n = make_network(R, C, D, Q, K, V, dO);
  memset(dV, 0, sizeof(float) * (size_t)C * D);
  if (threads <= 1) {
    float *matrixPRow = (float *)malloc(sizeof(float) * (size_t)C);
    for (int rowID = 0; rowID < R; ++rowID) {
      createMatrixPRow(&n, rowID, matrixPRow);
      for (int columnID = 0; columnID < C; ++columnID) {
        for (int d = 0; d < D; ++d) {
          size_t addressV = (size_t)columnID * D + d;
          size_t addressDerivativeO = (size_t)rowID * D + d;
          float dotProduct = dV[addressV];
          dotProduct += matrixPRow[columnID] * dO[addressDerivativeO];
          dV[addressV] = dotProduct;
        }
      }
    }
    free(matrixPRow);
    return;
  }
  float *partials = (float *)calloc((size_t)threads * C * D, sizeof(float));
#pragma omp parallel num_threads(threads)
  {
#ifdef _OPENMP
    int tid = omp_get_thread_num();
#else
    int tid = 0;
#endif
    float *mine = partials + (size_t)tid * C * D;
    float *matrixPRow = (float *)malloc(sizeof(float) * (size_t)C);
#pragma omp for schedule(static)
    for (int rowID = 0; rowID < R; ++rowID) {
      createMatrixPRow(&n, rowID, matrixPRow);
      for (int columnID = 0; columnID < C; ++columnID)
        for (int d = 0; d < D; ++d)
          mine[(size_t)columnID * D + d] += matrixPRow[columnID] * dO[(size_t)rowID * D + d];
    }
    free(matrixPRow);
  }
  for (int t = 0; t < threads; ++t)
    for (size_t i = 0; i < (size_t)C * D; ++i) dV[i] += partials[(size_t)t * C * D + i];
  free(partials);
}

/* Network.swift:352-372  derivativeK: dK = dS^T Q */
ORACLE_API void oracle_derivative_k(int R, int C, int D, const float *Q, const float *K,
                                    const float *V, const float *dO, float *dK, int threads) {
  Network n = make_network(R, C, D, Q, K, V, dO);
  memset(dK, 0, sizeof(float) * (size_t)C * D);
  int nt = threads > 1 ? threads : 1;
  float *partials = nt > 1 ? (float *)calloc((size_t)nt * C * D, sizeof(float)) : NULL;
#pragma omp parallel num_threads(nt) if (nt > 1)
  {
#ifdef _OPENMP
    int tid = omp_get_thread_num();
#else
    int tid = 0;
#endif
    float *mine = nt > 1 ? partials + (size_t)tid * C * D : dK;
    float *derivativeSRow = (float *)malloc(sizeof(float) * (size_t)C);
    float *scratchP = (float *)malloc(sizeof(float) * (size_t)C);
    float *scratchdP = (float *)malloc(sizeof(float) * (size_t)C);
    float *scratchO = (float *)malloc(sizeof(float) * (size_t)D);
#pragma omp for schedule(static)
    for (int rowID = 0; rowID < R; ++rowID) {
      createDerivativeSRow(&n, rowID, derivativeSRow, scratchP, scratchO, scratchdP);
      for (int columnID = 0; columnID < C; ++columnID) {
        for (int d = 0; d < D; ++d) {
          size_t addressK = (size_t)columnID * D + d;
          size_t addressQ = (size_t)rowID * D + d;
          float dotProduct = mine[addressK];
          dotProduct += derivativeSRow[columnID] * Q[addressQ];
          mine[addressK] = dotProduct;
        }
      }
    }
    free(derivativeSRow); free(scratchP); free(scratchdP); free(scratchO);
  }
  if (nt > 1) {
    for (int t = 0; t < nt; ++t)
      for (size_t i = 0; i < (size_t)C * D; ++i) dK[i] += partials[(size_t)t * C * D + i];
    free(partials);
  }
}

/* Network.swift:375-402  derivativeQ: dQ = dS K */
ORACLE_API void oracle_derivative_q(int R, int C, int D, const float *Q, const float *K,
                                    const float *V, const float *dO, float *dQ, int threads) {
  Network n = make_network(R, C, D, Q, K, V, dO);
#pragma omp parallel num_threads(threads > 1 ? threads : 1) if (threads > 1)
  {
    float *derivativeSRow = (float *)malloc(sizeof(float) * (size_t)C);
    float *scratchP = (float *)malloc(sizeof(float) * (size_t)C);
    float *scratchdP = (float *)malloc(sizeof(float) * (size_t)C);
    float *scratchO = (float *)malloc(sizeof(float) * (size_t)D);
#pragma omp for schedule(static)
    for (int rowID = 0; rowID < R; ++rowID) {
      createDerivativeSRow(&n, rowID, derivativeSRow, scratchP, scratchO, scratchdP);
      for (int d = 0; d < D; ++d) {
        float dotProduct = 0.0f;
        for (int columnID = 0; columnID < C; ++columnID) {
          dotProduct += derivativeSRow[columnID] * K[(size_t)columnID * D + d];
        }
        dQ[(size_t)rowID * D + d] = dotProduct;
      }
    }
    free(derivativeSRow); free(scratchP); free(scratchdP); free(scratchO);
  }
}

/* ------------------------------------------------------------------------------------------
 * 16-bit encode / decode exactly as the reference's test buffers do it
 * (Tests/FlashAttentionTests/Utilities/MTLContext+Buffers.swift:31-44, 66-76):
 *   FP16 = Float16(x) (IEEE round-to-nearest-even), BF16 = upper 16 bits (truncation).
 * ------------------------------------------------------------------------------------------ */
static uint16_t float_to_half_rne(float f) {
  uint32_t x; memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t mant = x & 0x007FFFFFu;
  int32_t exp = (int32_t)((x >> 23) & 0xFF);
  if (exp == 0xFF) return (uint16_t)(sign | 0x7C00u | (mant ? 0x200u : 0));
  int32_t e = exp - 127 + 15;
  if (e >= 0x1F) return (uint16_t)(sign | 0x7C00u);
  if (e <= 0) {
    if (e < -10) return (uint16_t)sign;
    mant |= 0x00800000u;
    uint32_t shift = (uint32_t)(14 - e);
    uint32_t half_mant = mant >> shift;
    uint32_t rem = mant & ((1u << shift) - 1);
    uint32_t halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (half_mant & 1))) half_mant++;
    return (uint16_t)(sign | half_mant);
  }
  uint32_t half = ((uint32_t)e << 10) | (mant >> 13);
  uint32_t rem = mant & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) half++;
  return (uint16_t)(sign | half);
}

static float half_to_float(uint16_t h) {
  uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1F;
  uint32_t mant = h & 0x3FFu;
  uint32_t x;
  if (exp == 0) {
    if (mant == 0) x = sign;
    else {
      int e = -1;
      do { mant <<= 1; e++; } while (!(mant & 0x400u));
      mant &= 0x3FFu;
      x = sign | ((uint32_t)(127 - 15 - e) << 23) | (mant << 13);
    }
  } else if (exp == 0x1F) {
    x = sign | 0x7F800000u | (mant << 13);
  } else {
    x = sign | ((exp - 15 + 127) << 23) | (mant << 13);
  }
  float f; memcpy(&f, &x, 4);
  return f;
}

/* precision: 0 = FP32, 1 = FP16, 2 = BF16 (GEMMOperandPrecision raw values) */
ORACLE_API void oracle_encode(const float *src, void *dst, size_t count, int precision) {
  if (precision == 0) { memcpy(dst, src, count * 4); return; }
  uint16_t *out = (uint16_t *)dst;
  for (size_t i = 0; i < count; ++i) {
    if (precision == 1) out[i] = float_to_half_rne(src[i]);
    else { uint32_t x; memcpy(&x, &src[i], 4); out[i] = (uint16_t)(x >> 16); }
  }
}

ORACLE_API void oracle_decode(const void *src, float *dst, size_t count, int precision) {
  if (precision == 0) { memcpy(dst, src, count * 4); return; }
  const uint16_t *in = (const uint16_t *)src;
  for (size_t i = 0; i < count; ++i) {
    if (precision == 1) dst[i] = half_to_float(in[i]);
    else { uint32_t x = (uint32_t)in[i] << 16; memcpy(&dst[i], &x, 4); }
  }
}

ORACLE_API int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
