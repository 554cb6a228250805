// fp32 CUDA-core (FFMA) tiled GEMM used for every dense contraction of the path
// that is not worth (or not safe for) the tensor-core route: tiny K (input
// layer, K = 2*13), tiny N (class layer, N = 9), the weight gradients, and as
// the exact-fp32 fallback for the hidden layers.
//
//   C(m,n) = post( sum_seg sum_k A_seg(m,k) * B_seg(k,n) )
//
// Up to two K segments ([h | ah*norm] against the two column blocks of W), so
// the concatenation of models.py:69-72 is never materialised.  Operands are
// described by their contiguous axis (A: K- or M-contiguous, B: K- or
// N-contiguous) so forward (X W^T), input gradient (dZ W) and weight gradient
// (dZ^T X, split over rows with a fixed-order reduction) share one kernel.
//
// Tiling: BM x BN x 16 block tile, 256 threads, TM x TN register tile, smem
// double buffer with register-staged global prefetch, 128-bit global loads
// when the operand is 16-byte aligned (zero-filled scalar loads otherwise).
#pragma once
#include "gte_common.cuh"

namespace gte {

constexpr int GEMM_THREADS = 256;
constexpr int GEMM_BK = 16;

struct GemmSeg {
  const float* A;
  int64_t lda;
  const float* B;
  int64_t ldb;
  int32_t K;
  int32_t vecA;  // 128-bit loads allowed for A / B of this segment
  int32_t vecB;
};

struct GemmArgs {
  GemmSeg seg[2];
  int32_t nseg;
  int32_t M, N;
  float* C;
  int64_t c_stride_m, c_stride_n;  // C(m,n) lives at C[m*c_stride_m + n*c_stride_n]
  const float* bias;               // per n, may be null
  const float* row_scale;          // per m, may be null
  int32_t accumulate;              // C += result
  int32_t vecC;                    // 128-bit stores allowed (c_stride_n == 1, 16-byte aligned rows)
  int32_t k_chunk;                 // split-K: rows of segment 0 per grid.z slice (multiple of 16); 0 = no split
  int64_t split_stride;            // floats between consecutive split partials
};

template <int BM, int BN, int TM, int TN, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(GEMM_THREADS) k_gemm_ffma(const GemmArgs g) {
  static_assert((BM / TM) * (BN / TN) == GEMM_THREADS, "thread tile mismatch");
  static_assert(TM == 8 || TM == 4, "TM");
  static_assert(TN == 8 || TN == 4 || TN == 1, "TN");
  constexpr int BK = GEMM_BK;
  constexpr int A_F4_TOTAL = BM * BK / 4;
  constexpr int B_F4_TOTAL = BN * BK / 4;
  constexpr int A_F4 = (A_F4_TOTAL + GEMM_THREADS - 1) / GEMM_THREADS;
  constexpr int B_F4 = (B_F4_TOTAL + GEMM_THREADS - 1) / GEMM_THREADS;
  constexpr int LDA_S = BM + 4;
  constexpr int LDB_S = BN + 4;

  __shared__ __align__(16) float As[2][BK][LDA_S];
  __shared__ __align__(16) float Bs[2][BK][LDB_S];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN);
  const int ty = tid / (BN / TN);
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // split-K range (segment 0 only)
  int32_t k_lo = 0, k_hi0 = g.seg[0].K;
  float* Cout = g.C;
  if (g.k_chunk > 0) {
    k_lo = blockIdx.z * g.k_chunk;
    k_hi0 = min(g.seg[0].K, k_lo + g.k_chunk);
    Cout += (int64_t)blockIdx.z * g.split_stride;
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[A_F4], rb[B_F4];

  auto load_A = [&](const GemmSeg& s, int32_t k0, int32_t kend) {
#pragma unroll
    for (int j = 0; j < A_F4; ++j) {
      const int i = tid + GEMM_THREADS * j;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (A_F4_TOTAL % GEMM_THREADS == 0 || i < A_F4_TOTAL) {
        if constexpr (A_KC) {
          const int row = i / (BK / 4), kq = i % (BK / 4);
          const int64_t m = m0 + row;
          const int32_t k = k0 + kq * 4;
          if (m < g.M && k < kend) {
            const float* p = s.A + m * s.lda + k;
            if (s.vecA && k + 3 < kend) {
              v = ldg4(p);
            } else {
              v.x = __ldg(p);
              if (k + 1 < kend) v.y = __ldg(p + 1);
              if (k + 2 < kend) v.z = __ldg(p + 2);
              if (k + 3 < kend) v.w = __ldg(p + 3);
            }
          }
        } else {
          const int krow = i / (BM / 4), mq = i % (BM / 4);
          const int32_t k = k0 + krow;
          const int64_t m = m0 + mq * 4;
          if (k < kend && m < g.M) {
            const float* p = s.A + (int64_t)k * s.lda + m;
            if (s.vecA && m + 3 < g.M) {
              v = ldg4(p);
            } else {
              v.x = __ldg(p);
              if (m + 1 < g.M) v.y = __ldg(p + 1);
              if (m + 2 < g.M) v.z = __ldg(p + 2);
              if (m + 3 < g.M) v.w = __ldg(p + 3);
            }
          }
        }
      }
      ra[j] = v;
    }
  };
  auto load_B = [&](const GemmSeg& s, int32_t k0, int32_t kend) {
#pragma unroll
    for (int j = 0; j < B_F4; ++j) {
      const int i = tid + GEMM_THREADS * j;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (B_F4_TOTAL % GEMM_THREADS == 0 || i < B_F4_TOTAL) {
        if constexpr (B_KC) {
          const int col = i / (BK / 4), kq = i % (BK / 4);
          const int n = n0 + col;
          const int32_t k = k0 + kq * 4;
          if (n < g.N && k < kend) {
            const float* p = s.B + (int64_t)n * s.ldb + k;
            if (s.vecB && k + 3 < kend) {
              v = ldg4(p);
            } else {
              v.x = __ldg(p);
              if (k + 1 < kend) v.y = __ldg(p + 1);
              if (k + 2 < kend) v.z = __ldg(p + 2);
              if (k + 3 < kend) v.w = __ldg(p + 3);
            }
          }
        } else {
          const int krow = i / (BN / 4), nq = i % (BN / 4);
          const int32_t k = k0 + krow;
          const int n = n0 + nq * 4;
          if (k < kend && n < g.N) {
            const float* p = s.B + (int64_t)k * s.ldb + n;
            if (s.vecB && n + 3 < g.N) {
              v = ldg4(p);
            } else {
              v.x = __ldg(p);
              if (n + 1 < g.N) v.y = __ldg(p + 1);
              if (n + 2 < g.N) v.z = __ldg(p + 2);
              if (n + 3 < g.N) v.w = __ldg(p + 3);
            }
          }
        }
      }
      rb[j] = v;
    }
  };
  auto store_AB = [&](int buf) {
#pragma unroll
    for (int j = 0; j < A_F4; ++j) {
      const int i = tid + GEMM_THREADS * j;
      if (A_F4_TOTAL % GEMM_THREADS == 0 || i < A_F4_TOTAL) {
        if constexpr (A_KC) {
          const int row = i / (BK / 4), kq = i % (BK / 4);
          As[buf][kq * 4 + 0][row] = ra[j].x;
          As[buf][kq * 4 + 1][row] = ra[j].y;
          As[buf][kq * 4 + 2][row] = ra[j].z;
          As[buf][kq * 4 + 3][row] = ra[j].w;
        } else {
          const int krow = i / (BM / 4), mq = i % (BM / 4);
          *reinterpret_cast<float4*>(&As[buf][krow][mq * 4]) = ra[j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < B_F4; ++j) {
      const int i = tid + GEMM_THREADS * j;
      if (B_F4_TOTAL % GEMM_THREADS == 0 || i < B_F4_TOTAL) {
        if constexpr (B_KC) {
          const int col = i / (BK / 4), kq = i % (BK / 4);
          Bs[buf][kq * 4 + 0][col] = rb[j].x;
          Bs[buf][kq * 4 + 1][col] = rb[j].y;
          Bs[buf][kq * 4 + 2][col] = rb[j].z;
          Bs[buf][kq * 4 + 3][col] = rb[j].w;
        } else {
          const int krow = i / (BN / 4), nq = i % (BN / 4);
          *reinterpret_cast<float4*>(&Bs[buf][krow][nq * 4]) = rb[j];
        }
      }
    }
  };

  // flattened tile walk over the segments
  int seg = 0;
  int32_t k0 = k_lo;
  int32_t kend = k_hi0;
  auto advance = [&]() {  // move (seg, k0) to the next tile; returns false when done
    k0 += BK;
    while (k0 >= kend) {
      ++seg;
      if (seg >= g.nseg) return false;
      k0 = 0;
      kend = g.seg[seg].K;
    }
    return true;
  };
  bool have = (k0 < kend);
  if (!have) {  // empty first segment slice: look for a later segment
    k0 = kend - BK;  // so that advance() steps past it
    have = advance();
  }
  if (have) {
    load_A(g.seg[seg], k0, kend);
    load_B(g.seg[seg], k0, kend);
    store_AB(0);
  }
  __syncthreads();
  int buf = 0;
  while (have) {
    const bool more = advance();
    if (more) {
      load_A(g.seg[seg], k0, kend);
      load_B(g.seg[seg], k0, kend);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
      if constexpr (TM == 8) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][BM / 2 + ty * 4]);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
        a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      } else {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      }
      if constexpr (TN == 8) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][BN / 2 + tx * 4]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
        b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
      } else if constexpr (TN == 4) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      } else {
        b[0] = Bs[buf][k][tx];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) store_AB(buf ^ 1);
    __syncthreads();
    buf ^= 1;
    have = more;
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int ml = (TM == 8) ? ((i < 4) ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4)) : ty * 4 + i;
    const int64_t m = m0 + ml;
    if (m >= g.M) continue;
    const float rs = g.row_scale ? __ldg(g.row_scale + m) : 1.0f;
    if constexpr (TN >= 4) {
#pragma unroll
      for (int jq = 0; jq < TN / 4; ++jq) {
        const int nl = (jq == 0) ? tx * 4 : BN / 2 + tx * 4;
        const int n = n0 + nl;
        if (n >= g.N) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[j] = acc[i][jq * 4 + j] * rs;
          if (g.bias && n + j < g.N) v[j] += __ldg(g.bias + n + j);
        }
        float* p = Cout + m * g.c_stride_m + (int64_t)n * g.c_stride_n;
        if (g.vecC && n + 3 < g.N) {
          float4 o = make_float4(v[0], v[1], v[2], v[3]);
          if (g.accumulate) {
            const float4 c = *reinterpret_cast<const float4*>(p);
            o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
          }
          *reinterpret_cast<float4*>(p) = o;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (n + j >= g.N) continue;
            float* pj = p + (int64_t)j * g.c_stride_n;
            *pj = g.accumulate ? v[j] + *pj : v[j];
          }
        }
      }
    } else {
      const int n = n0 + tx;
      if (n < g.N) {
        float v = acc[i][0] * rs;
        if (g.bias) v += __ldg(g.bias + n);
        float* p = Cout + m * g.c_stride_m + (int64_t)n * g.c_stride_n;
        if (g.accumulate) v += *p;
        *p = v;
      }
    }
  }
}

template <int BM, int BN, int TM, int TN, bool A_KC, bool B_KC>
static int launch_gemm_cfg(const GemmArgs& g, int splits, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div64(g.M, BM), (unsigned)ceil_div64(g.N, BN), (unsigned)(splits > 0 ? splits : 1));
  k_gemm_ffma<BM, BN, TM, TN, A_KC, B_KC><<<grid, GEMM_THREADS, 0, st>>>(g);
  GTE_CHECK_LAUNCH("k_gemm_ffma");
  return GTE_OK;
}

// pick the N tile by the output width
template <bool A_KC, bool B_KC>
static int launch_gemm(const GemmArgs& g, int splits, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return GTE_OK;
  if (g.N <= 16) return launch_gemm_cfg<128, 16, 8, 1, A_KC, B_KC>(g, splits, st);
  if (g.N <= 32) return launch_gemm_cfg<128, 32, 4, 4, A_KC, B_KC>(g, splits, st);
  if (g.N <= 64 || (g.N > 128 && g.N <= 192)) return launch_gemm_cfg<128, 64, 8, 4, A_KC, B_KC>(g, splits, st);
  return launch_gemm_cfg<128, 128, 8, 8, A_KC, B_KC>(g, splits, st);
}

}  // namespace gte
