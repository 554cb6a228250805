// Tall-skinny Gram reduction for the narrow weight gradients:
//
//     G[c][q] = sum_r P[r, c] * Q[r, q]        P: [n, wide]   Q = [Q1 | Q2 | 1]: [n, nq <= 32]
//
// e.g. the input layer dW = dz^T x with x only 13 columns wide, or the class
// layer with dz only 9 columns wide.  These are pure HBM streams over the wide
// operand (P is read exactly once; Q is tiny), so a GEMM tile is the wrong tool:
// here every thread owns one column of P, keeps NQ accumulators in registers,
// streams its column over a row chunk (coalesced 128-byte warp loads) and reads
// the Q rows as shared-memory broadcasts.  Row chunks are reduced afterwards in
// fixed order (deterministic, no atomics).  An implicit all-ones Q column gives
// the bias gradient (column sums of dz) in the same pass.
//
// Roofline: HBM, algorithmic bytes = 4*n*(wide + nq).
#pragma once
#include "gte_common.cuh"

namespace gte {

constexpr int GRAM_THREADS = 256;
constexpr int GRAM_ROWS = 16;  // Q rows staged per shared-memory tile (scalar kernel)

template <int NQ>
__global__ void __launch_bounds__(GRAM_THREADS)
    k_gram_tall(const float* __restrict__ P, int64_t ldp, int32_t wide, const float* __restrict__ Q1, int64_t ldq1,
                int32_t nq1, const float* __restrict__ Q2, int64_t ldq2, int32_t nq2, int32_t ones, int32_t n,
                int32_t rows_per_chunk, float* __restrict__ partial) {
  __shared__ __align__(16) float sq[GRAM_ROWS][NQ];
  const int c = blockIdx.y * GRAM_THREADS + threadIdx.x;
  const bool active = c < wide;
  const int32_t r0 = blockIdx.x * rows_per_chunk;
  const int32_t r1 = min(n, r0 + rows_per_chunk);
  const int nq = nq1 + nq2;
  float acc[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) acc[q] = 0.f;
  for (int32_t rb = r0; rb < r1; rb += GRAM_ROWS) {
    const int rows = min(GRAM_ROWS, r1 - rb);
    __syncthreads();
    for (int i = threadIdx.x; i < GRAM_ROWS * NQ; i += GRAM_THREADS) {
      const int rr = i / NQ, q = i % NQ;
      float v = 0.f;
      if (rr < rows) {
        const int64_t r = rb + rr;
        if (q < nq1) v = __ldg(Q1 + r * ldq1 + q);
        else if (q < nq) v = __ldg(Q2 + r * ldq2 + (q - nq1));
        else if (ones && q == nq) v = 1.0f;
      }
      sq[rr][q] = v;
    }
    __syncthreads();
    if (active) {
      float p[GRAM_ROWS];
#pragma unroll
      for (int rr = 0; rr < GRAM_ROWS; ++rr) p[rr] = (rr < rows) ? __ldg(P + (int64_t)(rb + rr) * ldp + c) : 0.f;
#pragma unroll
      for (int rr = 0; rr < GRAM_ROWS; ++rr) {
#pragma unroll
        for (int q4 = 0; q4 < NQ / 4; ++q4) {
          const float4 qv = *reinterpret_cast<const float4*>(&sq[rr][q4 * 4]);
          acc[q4 * 4 + 0] = fmaf(p[rr], qv.x, acc[q4 * 4 + 0]);
          acc[q4 * 4 + 1] = fmaf(p[rr], qv.y, acc[q4 * 4 + 1]);
          acc[q4 * 4 + 2] = fmaf(p[rr], qv.z, acc[q4 * 4 + 2]);
          acc[q4 * 4 + 3] = fmaf(p[rr], qv.w, acc[q4 * 4 + 3]);
        }
      }
    }
  }
  if (active) {
    float* out = partial + ((int64_t)blockIdx.x * wide + c) * NQ;
#pragma unroll
    for (int q4 = 0; q4 < NQ / 4; ++q4)
      *reinterpret_cast<float4*>(out + q4 * 4) = make_float4(acc[q4 * 4], acc[q4 * 4 + 1], acc[q4 * 4 + 2], acc[q4 * 4 + 3]);
  }
}

// 128-bit variant: a thread owns 4 adjacent columns of P (one LDG.128 per row) and NQ/QSPLIT of the
// q accumulators, so one shared-memory LDS.128 feeds 16 FMAs instead of 4.  Optionally also emits the
// column sums of Q1 (bias gradient when dz is the narrow operand).  partial layout unchanged.
template <int NQ, int QSPLIT>
__global__ void __launch_bounds__(64 * QSPLIT)
    k_gram_tall_v4(const float* __restrict__ P, int64_t ldp, int32_t wide, const float* __restrict__ Q1, int64_t ldq1,
                   int32_t nq1, const float* __restrict__ Q2, int64_t ldq2, int32_t nq2, int32_t ones, int32_t n,
                   int32_t rows_per_chunk, float* __restrict__ partial, float* __restrict__ qsum_partial) {
  constexpr int THREADS = 64 * QSPLIT;
  constexpr int QPT = NQ / QSPLIT;
  constexpr int TR = 16;  // rows per shared-memory tile
  static_assert(QPT % 4 == 0, "QPT");
  __shared__ __align__(16) float sq[TR][NQ];
  const int cg = threadIdx.x % 64;
  const int qh = threadIdx.x / 64;
  const int c = (blockIdx.y * 64 + cg) * 4;
  const bool active = c < wide;  // wide rounded up to 4 is readable (ldp % 4 == 0); extra columns are dropped at the end
  const int32_t r0 = blockIdx.x * rows_per_chunk;
  const int32_t r1 = min(n, r0 + rows_per_chunk);
  const int nq = nq1 + nq2;
  float acc[4][QPT];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int q = 0; q < QPT; ++q) acc[a][q] = 0.f;
  float qs = 0.f;  // column sum of Q (thread t < NQ owns column t)
  for (int32_t rb = r0; rb < r1; rb += TR) {
    const int rows = min(TR, r1 - rb);
    __syncthreads();
    for (int i = threadIdx.x; i < TR * NQ; i += THREADS) {
      const int rr = i / NQ, q = i % NQ;
      float v = 0.f;
      if (rr < rows) {
        const int64_t r = rb + rr;
        if (q < nq1) v = __ldg(Q1 + r * ldq1 + q);
        else if (q < nq) v = __ldg(Q2 + r * ldq2 + (q - nq1));
        else if (ones && q == nq) v = 1.0f;
      }
      sq[rr][q] = v;
    }
    float4 p[TR];
    if (active) {
#pragma unroll
      for (int rr = 0; rr < TR; ++rr)
        p[rr] = (rr < rows) ? __ldg(reinterpret_cast<const float4*>(P + (int64_t)(rb + rr) * ldp + c))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    if (qsum_partial && blockIdx.y == 0 && threadIdx.x < NQ) {
#pragma unroll
      for (int rr = 0; rr < TR; ++rr) qs += sq[rr][threadIdx.x];
    }
    if (active) {
#pragma unroll
      for (int rr = 0; rr < TR; ++rr) {
#pragma unroll
        for (int q4 = 0; q4 < QPT / 4; ++q4) {
          const float4 qv = *reinterpret_cast<const float4*>(&sq[rr][qh * QPT + q4 * 4]);
          const float pv[4] = {p[rr].x, p[rr].y, p[rr].z, p[rr].w};
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            acc[a][q4 * 4 + 0] = fmaf(pv[a], qv.x, acc[a][q4 * 4 + 0]);
            acc[a][q4 * 4 + 1] = fmaf(pv[a], qv.y, acc[a][q4 * 4 + 1]);
            acc[a][q4 * 4 + 2] = fmaf(pv[a], qv.z, acc[a][q4 * 4 + 2]);
            acc[a][q4 * 4 + 3] = fmaf(pv[a], qv.w, acc[a][q4 * 4 + 3]);
          }
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (c + a >= wide) continue;
      float* out = partial + ((int64_t)blockIdx.x * wide + c + a) * NQ + qh * QPT;
#pragma unroll
      for (int q4 = 0; q4 < QPT / 4; ++q4)
        *reinterpret_cast<float4*>(out + q4 * 4) =
            make_float4(acc[a][q4 * 4], acc[a][q4 * 4 + 1], acc[a][q4 * 4 + 2], acc[a][q4 * 4 + 3]);
    }
  }
  if (qsum_partial && blockIdx.y == 0 && threadIdx.x < NQ) qsum_partial[(int64_t)blockIdx.x * NQ + threadIdx.x] = qs;
}

// dst[q] (+)= sum over chunks of qsum_partial[chunk][q], q < nq1
__global__ void __launch_bounds__(RED_THREADS)
    k_gram_qsum_reduce(const float* __restrict__ qsum_partial, int chunks, int NQ, int32_t nq1, float* __restrict__ dst,
                       int accumulate) {
  __shared__ float red[RED_THREADS];
  const int i = threadIdx.x & 31;
  const bool valid = i < nq1;
  float s = reduce_partials_block(qsum_partial, chunks, NQ, i, valid, red);
  if ((threadIdx.x >> 5) != 0 || !valid) return;
  if (accumulate) s += dst[i];
  dst[i] = s;
}

// G[c][q] = sum over chunks (ascending), scattered to up to three destinations:
//   q <  nq1          -> D1[c*s1c + q*s1q]
//   nq1 <= q < nq     -> D2[c*s2c + (q-nq1)*s2q]
//   q == nq (ones)    -> Dones[c]
__global__ void __launch_bounds__(RED_THREADS)
    k_gram_reduce(const float* __restrict__ partial, int chunks, int32_t wide, int NQ, int32_t nq1, int32_t nq2,
                  float* __restrict__ D1, int64_t s1c, int64_t s1q, float* __restrict__ D2, int64_t s2c, int64_t s2q,
                  float* __restrict__ Dones, int accumulate) {
  __shared__ float red[RED_THREADS];
  const int64_t i = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
  const bool valid = i < (int64_t)wide * NQ;
  float s = reduce_partials_block(partial, chunks, (int64_t)wide * NQ, i, valid, red);
  if ((threadIdx.x >> 5) != 0 || !valid) return;
  const int c = (int)(i / NQ), q = (int)(i % NQ);
  const int nq = nq1 + nq2;
  float* dst;
  if (q < nq1) dst = D1 + c * s1c + q * s1q;
  else if (q < nq) dst = D2 + c * s2c + (q - nq1) * s2q;
  else if (q == nq && Dones) dst = Dones + c;
  else return;
  if (accumulate) s += *dst;
  *dst = s;
}

struct GramPlan {
  int NQ, chunks;
  int32_t rows_per_chunk;
  size_t ws_bytes;    // partial [chunks][wide][NQ] followed by qsum partial [chunks][NQ]
  size_t qsum_off;    // byte offset of the qsum partial inside the workspace
};

// shape-only plan (reproducible across devices)
static inline GramPlan gram_plan(int32_t n, int32_t wide, int32_t nq_total /* incl. ones column */) {
  GramPlan p;
  p.NQ = nq_total <= 8 ? 8 : (nq_total <= 16 ? 16 : (nq_total <= 24 ? 24 : 32));
  int64_t chunks = 1184 / ceil_div64(wide, 256);
  if (chunks < 1) chunks = 1;
  int64_t rpc = ceil_div64(n > 0 ? n : 1, chunks);
  if (rpc < 64) rpc = 64;
  rpc = ceil_div64(rpc, 16) * 16;
  p.rows_per_chunk = (int32_t)rpc;
  p.chunks = (int)ceil_div64(n > 0 ? n : 1, rpc);
  p.qsum_off = ((size_t)p.chunks * wide * p.NQ * 4 + 255) & ~size_t(255);
  p.ws_bytes = p.qsum_off + (((size_t)p.chunks * p.NQ * 4 + 255) & ~size_t(255));
  return p;
}

static inline bool gram_eligible(int32_t nq_total) { return nq_total <= 32; }

// Dq1sum (optional): column sums of Q1 (e.g. the bias gradient when dz is the narrow operand)
static inline int gram_tall(const float* P, int64_t ldp, int32_t wide, const float* Q1, int64_t ldq1, int32_t nq1,
                            const float* Q2, int64_t ldq2, int32_t nq2, bool ones, int32_t n, float* D1, int64_t s1c,
                            int64_t s1q, float* D2, int64_t s2c, int64_t s2q, float* Dones, float* Dq1sum,
                            int accumulate, float* ws, cudaStream_t st) {
  const GramPlan p = gram_plan(n, wide, nq1 + nq2 + (ones ? 1 : 0));
  float* qpart = Dq1sum ? reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + p.qsum_off) : nullptr;
  const bool vec = aligned16(P) && ldp % 4 == 0;
  const int o = ones ? 1 : 0;
  if (n > 0 && vec) {
    dim3 grid(p.chunks, (unsigned)ceil_div64(wide, 256));
#define GTE_GRAM_V4(NQV, QS) \
  k_gram_tall_v4<NQV, QS><<<grid, 64 * QS, 0, st>>>(P, ldp, wide, Q1, ldq1, nq1, Q2, ldq2, nq2, o, n, p.rows_per_chunk, ws, qpart)
    if (p.NQ == 8) GTE_GRAM_V4(8, 1);  // 8 q accumulators x 4 columns per thread everywhere
    else if (p.NQ == 16) GTE_GRAM_V4(16, 2);
    else if (p.NQ == 24) GTE_GRAM_V4(24, 3);
    else GTE_GRAM_V4(32, 4);
#undef GTE_GRAM_V4
    GTE_CHECK_LAUNCH("k_gram_tall_v4");
  } else if (n > 0) {
    dim3 grid(p.chunks, (unsigned)ceil_div64(wide, GRAM_THREADS));
    if (p.NQ == 8)
      k_gram_tall<8><<<grid, GRAM_THREADS, 0, st>>>(P, ldp, wide, Q1, ldq1, nq1, Q2, ldq2, nq2, o, n, p.rows_per_chunk, ws);
    else if (p.NQ == 16)
      k_gram_tall<16><<<grid, GRAM_THREADS, 0, st>>>(P, ldp, wide, Q1, ldq1, nq1, Q2, ldq2, nq2, o, n, p.rows_per_chunk, ws);
    else if (p.NQ == 24)
      k_gram_tall<24><<<grid, GRAM_THREADS, 0, st>>>(P, ldp, wide, Q1, ldq1, nq1, Q2, ldq2, nq2, o, n, p.rows_per_chunk, ws);
    else
      k_gram_tall<32><<<grid, GRAM_THREADS, 0, st>>>(P, ldp, wide, Q1, ldq1, nq1, Q2, ldq2, nq2, o, n, p.rows_per_chunk, ws);
    GTE_CHECK_LAUNCH("k_gram_tall");
  }
  const int64_t total = (int64_t)wide * p.NQ;
  k_gram_reduce<<<(unsigned)ceil_div64(total, 32), RED_THREADS, 0, st>>>(ws, n > 0 ? p.chunks : 0, wide, p.NQ, nq1, nq2, D1, s1c, s1q,
                                                                 D2, s2c, s2q, ones ? Dones : nullptr, accumulate);
  GTE_CHECK_LAUNCH("k_gram_reduce");
  if (Dq1sum) {
    if (n > 0 && !vec) return GTE_ERR_UNSUPPORTED;  // callers only request it on the vectorised route
    k_gram_qsum_reduce<<<1, RED_THREADS, 0, st>>>(qpart, n > 0 ? p.chunks : 0, p.NQ, nq1, Dq1sum, accumulate);
    GTE_CHECK_LAUNCH("k_gram_qsum_reduce");
  }
  return GTE_OK;
}

}  // namespace gte
