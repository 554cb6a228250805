// Streaming CUDA-core kernels for the NARROW dense operations of the model: the input layer
// (K = 2 * 13 input features) and the class layer (9 outputs).  These are `nn.Linear` and its autograd
// (/root/reference/src/components/graphs/models.py:27,63) at shapes where one operand is at most 32
// columns wide: a few thousand FMAs per node against ~1 KB of traffic per node, i.e. HBM streams, not
// GEMMs.  A tensor-core tile pipeline (TMA -> 3xTF32 split -> tcgen05.mma -> TMEM -> epilogue) pays its
// full per-tile overhead for almost no arithmetic there; the kernels below stream the wide operand once
// at full width and keep the narrow one in shared memory / registers.  Exact fp32 FMA arithmetic.
//
//   k_gram_stream : C[a][b] = sum_n P[n,a] * Q[n,b]      P wide (<= 256), Q = [Q1|Q2] narrow (<= 32)
//                   = the weight gradients  dW1 = dz^T [h|ah]  and  dW3 = [dz|G]^T y2   (+ column sums of Q1)
//   k_wide_out    : out[n,c] = sum_j A[n,j] * B[j][c] (+ bias) (+ LayerNorm + ReLU)    A = [A1|A2] narrow, c wide
//                   = the input-layer forward (fused bias + LayerNorm + ReLU, models.py:63-66) and the
//                     class-layer input gradient  dy2 = dz Ws + G Wn
//
// Roofline: HBM.  Algorithmic bytes: gram 4*n*(wide + narrow); wide_out 4*n*(narrow + wide) (x2 outputs with LN).
#include "gte_common.cuh"

namespace gte {

// ---------------------------------------------------------------------------------------------------
// Tall-skinny Gram reduction.  CTA = 256 threads = 64 column chunks of P (4 columns each) x 4 groups of
// the Q columns (NB <= 8 each); a thread keeps its 4 x NB block of C in registers while the CTA streams a
// contiguous range of rows through a double-buffered shared-memory tile (P rows by 16-byte cp.async, Q
// values by 4-byte cp.async into 8-float group slots).  Row ranges are reduced afterwards in fixed order.
constexpr int GS_THREADS = 256;
constexpr int GS_TR = 32;      // rows per tile
constexpr int GS_QSLOT = 32;   // floats per Q row in shared memory: 4 groups x 8

template <int NB>
__global__ void __launch_bounds__(GS_THREADS)
    k_gram_stream(const float* __restrict__ P, int64_t ldp, int32_t wide, const float* __restrict__ Q1, int64_t ldq1,
                  int32_t nq1, const float* __restrict__ Q2, int64_t ldq2, int32_t nq2, int32_t n, int32_t rows_per_cta,
                  float* __restrict__ partial) {
  extern __shared__ __align__(16) float gs_smem[];
  const int ac_n = (wide + 3) / 4;                     // 16-byte chunks per P row
  const int p_pitch = ac_n * 4;                        // floats
  float* sP[2] = {gs_smem, gs_smem + GS_TR * p_pitch};
  float* sQ[2] = {gs_smem + 2 * GS_TR * p_pitch, gs_smem + 2 * GS_TR * p_pitch + GS_TR * GS_QSLOT};
  const int tid = threadIdx.x;
  const int ac = tid & 63, bg = tid >> 6;
  const int nq = nq1 + nq2;
  const int32_t r0 = blockIdx.x * rows_per_cta;
  const int32_t r1 = min(n, r0 + rows_per_cta);
  // zero the Q slots once: padding slots are never written again
  for (int i = tid; i < 2 * GS_TR * GS_QSLOT; i += GS_THREADS) sQ[0][i] = 0.f;
  __syncthreads();

  auto stage = [&](int32_t rb, int s) {
    const int rows = min(GS_TR, r1 - rb);
    // index math without divisions: 64 chunk slots per P row, 32 value slots per Q row
    for (int i = tid; i < GS_TR * 64; i += GS_THREADS) {
      const int rr = i >> 6, c4 = i & 63;
      if (c4 >= ac_n) continue;
      float* dst = sP[s] + rr * p_pitch + c4 * 4;
      if (rr < rows) {
        const float* src = P + (int64_t)(rb + rr) * ldp + c4 * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                     : "memory");
      } else {
        *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    for (int i = tid; i < GS_TR * GS_QSLOT; i += GS_THREADS) {
      const int rr = i >> 5, slot = i & 31;
      const int j = slot & 7, q = (slot >> 3) * NB + j;
      if (j >= NB || q >= nq) continue;
      float* dst = sQ[s] + rr * GS_QSLOT + slot;
      if (rr < rows) {
        const float* src = q < nq1 ? Q1 + (int64_t)(rb + rr) * ldq1 + q : Q2 + (int64_t)(rb + rr) * ldq2 + (q - nq1);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                     : "memory");
      } else {
        *dst = 0.f;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[4][NB];
  float qsum[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    qsum[j] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][j] = 0.f;
  }
  const bool active = ac < ac_n;
  int s = 0;
  if (r0 < r1) stage(r0, 0);
  for (int32_t rb = r0; rb < r1; rb += GS_TR, s ^= 1) {
    if (rb + GS_TR < r1) {
      stage(rb + GS_TR, s ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (active) {
      const float* p = sP[s] + ac * 4;
      const float* q = sQ[s] + bg * 8;
#pragma unroll 4
      for (int rr = 0; rr < GS_TR; ++rr) {
        const float4 pv = *reinterpret_cast<const float4*>(p + rr * p_pitch);
        float qv[8];
        *reinterpret_cast<float4*>(qv) = *reinterpret_cast<const float4*>(q + rr * GS_QSLOT);
        if (NB > 4) *reinterpret_cast<float4*>(qv + 4) = *reinterpret_cast<const float4*>(q + rr * GS_QSLOT + 4);
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          acc[0][j] = fmaf(pv.x, qv[j], acc[0][j]);
          acc[1][j] = fmaf(pv.y, qv[j], acc[1][j]);
          acc[2][j] = fmaf(pv.z, qv[j], acc[2][j]);
          acc[3][j] = fmaf(pv.w, qv[j], acc[3][j]);
          if (ac == 0) qsum[j] += qv[j];
        }
      }
    }
    __syncthreads();  // the tile just read is the one the next iteration's stage() overwrites
  }
  // partial[cta][a][slot], slot = bg*8 + j; row `wide` holds the column sums of Q
  float* out = partial + (int64_t)blockIdx.x * (int64_t)(wide + 1) * GS_QSLOT;
  if (active) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int a = ac * 4 + i;
      if (a < wide)
#pragma unroll
        for (int j = 0; j < NB; ++j) out[(int64_t)a * GS_QSLOT + bg * 8 + j] = acc[i][j];
    }
    if (ac == 0)
#pragma unroll
      for (int j = 0; j < NB; ++j) out[(int64_t)wide * GS_QSLOT + bg * 8 + j] = qsum[j];
  }
}

// C (+)= fixed-order sum of the per-CTA partials, scattered to the two destination blocks
//   (a, b) with b <  nq1 -> out1[a*sa1 + b*sb1];  b >= nq1 -> out2[a*sa2 + (b-nq1)*sb2];  a == wide -> qsum[b] (b < nq1)
template <int NB>
__global__ void __launch_bounds__(RED_THREADS)
    k_gram_stream_reduce(const float* __restrict__ partial, int nparts, int32_t wide, int32_t nq1, int32_t nq2,
                         float* __restrict__ out1, int64_t sa1, int64_t sb1, float* __restrict__ out2, int64_t sa2,
                         int64_t sb2, float* __restrict__ qsum, int accumulate) {
  __shared__ float red[RED_THREADS];
  const int64_t stride = (int64_t)(wide + 1) * GS_QSLOT;
  const int64_t i = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
  const int a = (int)(i / GS_QSLOT), slot = (int)(i % GS_QSLOT);
  const int b = (slot / 8) * NB + (slot % 8);
  const bool valid = i < stride && (slot % 8) < NB && b < nq1 + nq2 && (a < wide || (qsum != nullptr && b < nq1));
  float s = reduce_partials_block(partial, nparts, stride, i, valid, red);
  if ((threadIdx.x >> 5) != 0 || !valid) return;
  float* dst;
  if (a == wide) dst = qsum + b;
  else if (b < nq1) dst = out1 + a * sa1 + b * sb1;
  else dst = out2 + a * sa2 + (b - nq1) * sb2;
  if (accumulate) s += *dst;
  *dst = s;
}

static int gs_nb(int nq) {  // Q columns per thread group (4 groups)
  return (nq + 3) / 4;
}
static int gs_grid(int32_t n, int32_t* rows_per_cta) {
  int64_t target = (int64_t)sm_count() * 3;
  int64_t rows = ceil_div64(ceil_div64(n, target), GS_TR) * GS_TR;
  if (rows < GS_TR) rows = GS_TR;
  *rows_per_cta = (int32_t)rows;
  return (int)ceil_div64(n, rows);
}
static size_t gs_smem_bytes(int wide) {
  const int p_pitch = (wide + 3) / 4 * 4;
  return (size_t)(2 * GS_TR * p_pitch + 2 * GS_TR * GS_QSLOT) * 4;
}

// ---------------------------------------------------------------------------------------------------
// Narrow-contraction, wide-output product with an optional fused bias + LayerNorm + ReLU epilogue.
// A warp owns 8 rows at a time; lane l owns output columns l, l+32, ... (NI <= 8 of them), so a row's
// outputs live in the warp's registers: LayerNorm statistics are two warp reductions over data that is
// already there (true two-pass mean / variance), and every store is a coalesced 128-byte line.  The
// [J][C] matrix sits in shared memory (lanes read consecutive floats), the 8 narrow rows in a per-warp
// shared buffer read as broadcasts.
constexpr int WO_THREADS = 256;
constexpr int WO_ROWS = 8;
constexpr int WO_JP = 32;  // shared slots per narrow row: segment 1 at 0..15, segment 2 at 16..31

template <int NI, bool LN>
__global__ void __launch_bounds__(WO_THREADS)
    k_wide_out(const float* __restrict__ A1, int64_t lda1, int32_t k1, const float* __restrict__ A2, int64_t lda2,
               int32_t k2, const float* __restrict__ B1 /* element (j, c) at j*sj + c*sc */, const float* __restrict__ B2,
               int64_t sj, int64_t sc, const float* __restrict__ bias,
               const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int relu,
               const float* __restrict__ row_scale, float* __restrict__ z, int64_t ldz, float* __restrict__ y, int64_t ldy,
               float* __restrict__ mean_out, float* __restrict__ rstd_out, int32_t n, int32_t C) {
  extern __shared__ __align__(16) float wo_smem[];
  const int cpad = NI * 32;
  float* sB = wo_smem;                                   // [WO_JP][cpad]
  float* sA = wo_smem + WO_JP * cpad;                    // [warps][WO_ROWS][WO_JP]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < WO_JP * cpad; i += WO_THREADS) {
    const int j = i / cpad, c = i - j * cpad;
    float v = 0.f;
    if (c < C) {
      if (j < 16) { if (j < k1) v = __ldg(B1 + (int64_t)j * sj + (int64_t)c * sc); }
      else if (j - 16 < k2) v = __ldg(B2 + (int64_t)(j - 16) * sj + (int64_t)c * sc);
    }
    sB[i] = v;
  }
  __syncthreads();
  float* myA = sA + warp * (WO_ROWS * WO_JP);
  const int ch1 = (k1 + 3) / 4, ch2 = (k2 + 3) / 4;
  float bia[NI], gam[NI], bet[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int c = lane + 32 * i;
    bia[i] = (bias && c < C) ? __ldg(bias + c) : 0.f;
    gam[i] = (LN && c < C) ? __ldg(gamma + c) : 0.f;
    bet[i] = (LN && c < C) ? __ldg(beta + c) : 0.f;
  }
  const int64_t ngroups = ((int64_t)n + WO_ROWS - 1) / WO_ROWS;
  for (int64_t grp = (int64_t)blockIdx.x * (WO_THREADS / 32) + warp; grp < ngroups; grp += (int64_t)gridDim.x * (WO_THREADS / 32)) {
    const int64_t row0 = grp * WO_ROWS;
    // stage the 8 narrow rows (columns >= k are padding: zeroed, they may hold anything)
    __syncwarp();
    for (int i = lane; i < WO_ROWS * 8; i += 32) {  // 8 x (4 chunks seg 1 + 4 chunks seg 2)
      const int r = i >> 3, c4 = i & 7;
      const int seg = c4 >> 2, cc = c4 & 3;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int64_t row = row0 + r;
      if (row < n) {
        if (seg == 0 && cc < ch1) v = __ldg(reinterpret_cast<const float4*>(A1 + row * lda1 + cc * 4));
        if (seg == 1 && cc < ch2) v = __ldg(reinterpret_cast<const float4*>(A2 + row * lda2 + cc * 4));
        const int kk = seg == 0 ? k1 : k2;
        if (cc * 4 + 0 >= kk) v.x = 0.f;
        if (cc * 4 + 1 >= kk) v.y = 0.f;
        if (cc * 4 + 2 >= kk) v.z = 0.f;
        if (cc * 4 + 3 >= kk) v.w = 0.f;
      }
      *reinterpret_cast<float4*>(myA + r * WO_JP + c4 * 4) = v;
    }
    __syncwarp();
    float acc[WO_ROWS][NI];
#pragma unroll
    for (int r = 0; r < WO_ROWS; ++r)
#pragma unroll
      for (int i = 0; i < NI; ++i) acc[r][i] = bia[i];
#pragma unroll 1
    for (int seg = 0; seg < 2; ++seg) {
      const int nch = seg == 0 ? ch1 : ch2;
      for (int c4 = 0; c4 < nch; ++c4) {
        const int j0 = seg * 16 + c4 * 4;
        float4 av[WO_ROWS];
#pragma unroll
        for (int r = 0; r < WO_ROWS; ++r) av[r] = *reinterpret_cast<const float4*>(myA + r * WO_JP + j0);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float w[NI];
#pragma unroll
          for (int i = 0; i < NI; ++i) w[i] = sB[(j0 + jj) * cpad + lane + 32 * i];
#pragma unroll
          for (int r = 0; r < WO_ROWS; ++r) {
            const float a = jj == 0 ? av[r].x : jj == 1 ? av[r].y : jj == 2 ? av[r].z : av[r].w;
#pragma unroll
            for (int i = 0; i < NI; ++i) acc[r][i] = fmaf(a, w[i], acc[r][i]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < WO_ROWS; ++r) {
      const int64_t row = row0 + r;
      if (row >= n) break;  // warp-uniform
      if (row_scale) {
        const float sc_ = __ldg(row_scale + row);
#pragma unroll
        for (int i = 0; i < NI; ++i) acc[r][i] *= sc_;
      }
      float* zr = z + row * ldz;
#pragma unroll
      for (int i = 0; i < NI; ++i)
        if (lane + 32 * i < C) zr[lane + 32 * i] = acc[r][i];
      if (LN) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NI; ++i)
          if (lane + 32 * i < C) s += acc[r][i];
        const float mu = warp_sum(s) / (float)C;
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < NI; ++i)
          if (lane + 32 * i < C) {
            const float d = acc[r][i] - mu;
            v = fmaf(d, d, v);
          }
        const float rs = 1.0f / sqrtf(warp_sum(v) / (float)C + eps);
        float* yr = y + row * ldy;
#pragma unroll
        for (int i = 0; i < NI; ++i)
          if (lane + 32 * i < C) {
            float o = (acc[r][i] - mu) * rs * gam[i] + bet[i];
            if (relu) o = fmaxf(o, 0.f);
            yr[lane + 32 * i] = o;
          }
        if (lane == 0) {
          mean_out[row] = mu;
          rstd_out[row] = rs;
        }
      }
    }
  }
}

static size_t wo_smem_bytes(int ni) { return (size_t)(WO_JP * ni * 32 + (WO_THREADS / 32) * WO_ROWS * WO_JP) * 4; }

}  // namespace gte

using namespace gte;

extern "C" size_t gte_gram_stream_workspace_bytes(int32_t n, int32_t wide) {
  if (n <= 0 || wide <= 0) return 256;
  int32_t rows;
  const int grid = gs_grid(n, &rows);
  return (size_t)grid * (size_t)(wide + 1) * GS_QSLOT * 4 + 256;
}

extern "C" int gte_gram_stream(const float* P, int64_t ldp, int32_t wide, const float* Q1, int64_t ldq1, int32_t nq1,
                               const float* Q2, int64_t ldq2, int32_t nq2, int32_t n, float* out1, int64_t sa1,
                               int64_t sb1, float* out2, int64_t sa2, int64_t sb2, float* qsum, int accumulate, void* ws,
                               size_t ws_bytes, gte_stream_t stream) {
  GTE_CHECK_ARG(n >= 0 && wide > 0 && nq1 > 0 && nq2 >= 0, "gte_gram_stream: bad size");
  if (wide > 256 || nq1 + nq2 > 32)
    return fail(GTE_ERR_UNSUPPORTED, "gte_gram_stream: wide=%d (<= 256) nq=%d (<= 32) not supported", wide, nq1 + nq2);
  GTE_CHECK_ARG(P && Q1 && out1 && (nq2 == 0 || (Q2 && out2)), "gte_gram_stream: null argument");
  GTE_CHECK_ARG(aligned16(P) && ldp % 4 == 0 && ldp >= (wide + 3) / 4 * 4, "gte_gram_stream: P rows must be 16-byte aligned and padded to 4");
  GTE_CHECK_ARG(ldq1 >= nq1 && (nq2 == 0 || ldq2 >= nq2), "gte_gram_stream: leading dimension < columns");
  if (ws_bytes < gte_gram_stream_workspace_bytes(n, wide) || !ws)
    return fail(GTE_ERR_WORKSPACE, "gte_gram_stream: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int nq = nq1 + nq2;
  const int nb = gs_nb(nq);
  int32_t rows = GS_TR;
  const int grid = n > 0 ? gs_grid(n, &rows) : 0;
  float* partial = static_cast<float*>(ws);
  const size_t smem = gs_smem_bytes(wide);
#define GTE_GS_GO(NBV)                                                                                                    \
  do {                                                                                                                    \
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(&k_gram_stream<NBV>), smem, "k_gram_stream")) return rc; \
    if (grid > 0) {                                                                                                       \
      k_gram_stream<NBV><<<grid, GS_THREADS, smem, st>>>(P, ldp, wide, Q1, ldq1, nq1, Q2, ldq2, nq2, n, rows, partial);  \
      GTE_CHECK_LAUNCH("k_gram_stream");                                                                                  \
    }                                                                                                                     \
    const int64_t total = (int64_t)(wide + 1) * GS_QSLOT;                                                                 \
    k_gram_stream_reduce<NBV><<<(unsigned)ceil_div64(total, 32), RED_THREADS, 0, st>>>(                                   \
        partial, grid, wide, nq1, nq2, out1, sa1, sb1, out2, sa2, sb2, qsum, accumulate);                                 \
    GTE_CHECK_LAUNCH("k_gram_stream_reduce");                                                                             \
    return GTE_OK;                                                                                                        \
  } while (0)
  switch (nb) {
    case 1: GTE_GS_GO(1);
    case 2: GTE_GS_GO(2);
    case 3: GTE_GS_GO(3);
    case 4: GTE_GS_GO(4);
    case 5: GTE_GS_GO(5);
    case 6: GTE_GS_GO(6);
    case 7: GTE_GS_GO(7);
    default: GTE_GS_GO(8);
  }
#undef GTE_GS_GO
}

extern "C" int gte_wide_out(const float* A1, int64_t lda1, int32_t k1, const float* A2, int64_t lda2, int32_t k2,
                            const float* B1, const float* B2, int64_t sj, int64_t sc, const float* bias, const float* gamma,
                            const float* beta, float eps, int relu, int fuse_ln, const float* row_scale, float* z,
                            int64_t ldz, float* y, int64_t ldy, float* mean, float* rstd, int32_t n, int32_t C,
                            gte_stream_t stream) {
  GTE_CHECK_ARG(n >= 0 && C > 0 && k1 > 0 && k2 >= 0, "gte_wide_out: bad size");
  if (C > 256 || k1 > 16 || k2 > 16)
    return fail(GTE_ERR_UNSUPPORTED, "gte_wide_out: C=%d (<= 256), k1=%d k2=%d (<= 16 each) not supported", C, k1, k2);
  if (n == 0) return GTE_OK;
  GTE_CHECK_ARG(A1 && B1 && z && ldz >= C && (k2 == 0 || (A2 && B2)), "gte_wide_out: null argument");
  GTE_CHECK_ARG(aligned16(A1) && lda1 % 4 == 0 && lda1 >= (k1 + 3) / 4 * 4 &&
                    (k2 == 0 || (aligned16(A2) && lda2 % 4 == 0 && lda2 >= (k2 + 3) / 4 * 4)),
                "gte_wide_out: narrow operands must have 16-byte aligned rows padded to 4 columns");
  GTE_CHECK_ARG(!fuse_ln || (gamma && beta && y && mean && rstd && ldy >= C), "gte_wide_out: LayerNorm needs gamma, beta, y, mean, rstd");
  cudaStream_t st = as_stream(stream);
  const int ni = (C + 31) / 32;
  const size_t smem = wo_smem_bytes(ni);
  const int64_t ngroups = ceil_div64(n, WO_ROWS);
  int64_t grid = (int64_t)sm_count() * 2;
  const int64_t need = ceil_div64(ngroups, WO_THREADS / 32);
  if (grid > need) grid = need;
#define GTE_WO_GO(NIV)                                                                                                  \
  do {                                                                                                                  \
    if (int rc = ensure_dynamic_smem(fuse_ln ? reinterpret_cast<const void*>(&k_wide_out<NIV, true>)                    \
                                             : reinterpret_cast<const void*>(&k_wide_out<NIV, false>),                  \
                                     smem, "k_wide_out"))                                                               \
      return rc;                                                                                                        \
    if (fuse_ln)                                                                                                        \
      k_wide_out<NIV, true><<<(unsigned)grid, WO_THREADS, smem, st>>>(A1, lda1, k1, A2, lda2, k2, B1, B2, sj, sc, bias, gamma, \
                                                                      beta, eps, relu, row_scale, z, ldz, y, ldy, mean, rstd, n, C); \
    else                                                                                                                \
      k_wide_out<NIV, false><<<(unsigned)grid, WO_THREADS, smem, st>>>(A1, lda1, k1, A2, lda2, k2, B1, B2, sj, sc, bias, gamma, \
                                                                       beta, eps, relu, row_scale, z, ldz, y, ldy, mean, rstd, n, C); \
    GTE_CHECK_LAUNCH("k_wide_out");                                                                                     \
    return GTE_OK;                                                                                                      \
  } while (0)
  switch (ni) {
    case 1: GTE_WO_GO(1);
    case 2: GTE_WO_GO(2);
    case 3: GTE_WO_GO(3);
    case 4: GTE_WO_GO(4);
    case 5: GTE_WO_GO(5);
    case 6: GTE_WO_GO(6);
    case 7: GTE_WO_GO(7);
    default: GTE_WO_GO(8);
  }
#undef GTE_WO_GO
}
