// Gather + segment-reduce SpMM over compressed rows: the neighbour aggregation
// of the GraphSAGE layers, forward (CSC) and backward (transposed CSR).
//
// Replaces `g.update_all(fn.u_mul_e('h','feat','m'), fn.sum('m','h'))` and the
// `ah * norm` of `concat` (/root/reference/src/components/graphs/models.py:53-54,
// 69-71), `fn.mean` (models.py:149), and autograd's reverse-graph SpMM.
//
// Kernel `k_spmm_rows`: a group of G lanes owns one output row (G = 4..32 by
// feature width, so narrow rows pack 8 per warp); the group's lanes fetch up to
// G (index, weight) pairs with one coalesced load and broadcast them by
// shuffle; each lane accumulates C 128-bit column chunks in registers in edge
// order (deterministic, no atomics); the edge weight, the optional source-side
// scale (backward: norm[dst]), the degree normalisation and the optional addend
// (backward: the self-path gradient) are fused.
//
// Roofline: HBM.  Algorithmic bytes per launch = 8*N*F + 8*E + 4*N  (read x,
// write y, indptr, indices, weights); the gathered rows
// (4*E*F bytes) are served by L1/L2 because a page's sources are page-local.
#include "gte_common.cuh"

namespace gte {

template <int VEC>
struct Acc;
template <>
struct Acc<4> {
  float4 v;
  __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void fma(float w, const float4& x) {
    v.x = fmaf(w, x.x, v.x);
    v.y = fmaf(w, x.y, v.y);
    v.z = fmaf(w, x.z, v.z);
    v.w = fmaf(w, x.w, v.w);
  }
};
template <>
struct Acc<1> {
  float v;
  __device__ __forceinline__ void zero() { v = 0.f; }
  __device__ __forceinline__ void fma(float w, float x) { v = fmaf(w, x, v); }
};

template <int VEC>
struct VecT;
template <>
struct VecT<4> {
  using type = float4;
  static __device__ __forceinline__ float4 load(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
};
template <>
struct VecT<1> {
  using type = float;
  static __device__ __forceinline__ float load(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ float zero() { return 0.f; }
};

constexpr int SPMM_THREADS = 256;
constexpr int EDGE_UNROLL = 4;

template <int VEC, int G, int C>
__global__ void __launch_bounds__(SPMM_THREADS)
    k_spmm_rows(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ w,
                const float* __restrict__ pre_scale, const float* __restrict__ row_norm, int mode,
                const float* __restrict__ x, int64_t ldx, const float* __restrict__ addend, int64_t ldadd,
                float* __restrict__ y, int64_t ldy, int32_t n_rows, int32_t f) {
  using V = typename VecT<VEC>::type;
  constexpr int ROWS_PER_BLOCK = SPMM_THREADS / G;
  const int lane = threadIdx.x % G;
  const int grp = threadIdx.x / G;
  const int64_t row = (int64_t)blockIdx.x * ROWS_PER_BLOCK + grp;
  const int col_base = blockIdx.y * (G * C * VEC);
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
  if (row >= n_rows) return;  // whole group leaves together; shuffles below use the group mask

  const int32_t beg = indptr[row], end = indptr[row + 1];
  Acc<VEC> acc[C];
  int col[C];
  bool on[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    acc[c].zero();
    col[c] = col_base + (lane + c * G) * VEC;
    on[c] = col[c] < f;
  }

  for (int32_t base = beg; base < end; base += G) {
    const int32_t mine = base + lane;
    int32_t idx = 0;
    float wv = 0.f;
    if (mine < end) {
      idx = __ldg(indices + mine);
      wv = w ? __ldg(w + mine) : 1.0f;
      if (pre_scale) wv *= __ldg(pre_scale + idx);
    }
    const int cnt = min(G, end - base);
    int t = 0;
    for (; t + EDGE_UNROLL <= cnt; t += EDGE_UNROLL) {
      V xv[EDGE_UNROLL][C];
      float wt[EDGE_UNROLL];
#pragma unroll
      for (int u = 0; u < EDGE_UNROLL; ++u) {
        const int32_t src = __shfl_sync(gmask, idx, t + u, G);
        wt[u] = __shfl_sync(gmask, wv, t + u, G);
        const float* xr = x + (int64_t)src * ldx;
#pragma unroll
        for (int c = 0; c < C; ++c) xv[u][c] = on[c] ? VecT<VEC>::load(xr + col[c]) : VecT<VEC>::zero();
      }
#pragma unroll
      for (int u = 0; u < EDGE_UNROLL; ++u)
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c].fma(wt[u], xv[u][c]);
    }
    for (; t < cnt; ++t) {
      const int32_t src = __shfl_sync(gmask, idx, t, G);
      const float wt = __shfl_sync(gmask, wv, t, G);
      const float* xr = x + (int64_t)src * ldx;
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (on[c]) acc[c].fma(wt, VecT<VEC>::load(xr + col[c]));
    }
  }

  float post = 1.0f;
  bool divide = false;
  if (mode == GTE_AGG_SUM_NORM) {
    post = __ldg(row_norm + row);
  } else if (mode == GTE_AGG_MEAN) {
    const int32_t deg = end - beg;
    post = (float)(deg > 1 ? deg : 1);
    divide = true;
  }
  float* yr = y + row * ldy;
  const float* ar = addend ? addend + row * ldadd : nullptr;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    if (!on[c]) continue;
    if constexpr (VEC == 4) {
      float4 r = acc[c].v;
      if (divide) {
        r.x /= post; r.y /= post; r.z /= post; r.w /= post;
      } else {
        r.x *= post; r.y *= post; r.z *= post; r.w *= post;
      }
      if (ar) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(ar + col[c]));
        r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
      }
      *reinterpret_cast<float4*>(yr + col[c]) = r;
    } else {
      float r = divide ? acc[c].v / post : acc[c].v * post;
      if (ar) r += __ldg(ar + col[c]);
      yr[col[c]] = r;
    }
  }
}

template <int VEC, int G, int C>
static int launch_spmm(const int32_t* indptr, const int32_t* indices, const float* w, const float* pre_scale,
                       const float* row_norm, int mode, const float* x, int64_t ldx, const float* addend,
                       int64_t ldadd, float* y, int64_t ldy, int32_t n_rows, int32_t f, cudaStream_t st) {
  constexpr int ROWS_PER_BLOCK = SPMM_THREADS / G;
  const int cols_per_block = G * C * VEC;
  dim3 grid((unsigned)ceil_div64(n_rows, ROWS_PER_BLOCK), (unsigned)ceil_div64(f, cols_per_block));
  k_spmm_rows<VEC, G, C><<<grid, SPMM_THREADS, 0, st>>>(indptr, indices, w, pre_scale, row_norm, mode, x, ldx, addend,
                                                       ldadd, y, ldy, n_rows, f);
  GTE_CHECK_LAUNCH("k_spmm_rows");
  return GTE_OK;
}

}  // namespace gte

using namespace gte;

extern "C" int gte_spmm(const int32_t* indptr, const int32_t* indices, const float* w, const float* pre_scale,
                        const float* row_norm, int mode, const float* x, int64_t ldx, const float* addend,
                        int64_t ldadd, float* y, int64_t ldy, int32_t n_rows, int32_t f, gte_stream_t stream) {
  GTE_CHECK_ARG(n_rows >= 0 && f >= 0, "gte_spmm: negative size");
  GTE_CHECK_ARG(mode == GTE_AGG_SUM || mode == GTE_AGG_SUM_NORM || mode == GTE_AGG_MEAN, "gte_spmm: bad mode %d", mode);
  if (n_rows == 0 || f == 0) return GTE_OK;
  GTE_CHECK_ARG(indptr && indices && x && y, "gte_spmm: null argument");
  GTE_CHECK_ARG(mode != GTE_AGG_SUM_NORM || row_norm, "gte_spmm: SUM_NORM needs row_norm");
  GTE_CHECK_ARG(ldx >= f && ldy >= f && (!addend || ldadd >= f), "gte_spmm: leading dimension < f");
  GTE_CHECK_ARG(x != y, "gte_spmm: x and y must not alias");
  cudaStream_t st = as_stream(stream);
  const bool vec = aligned16(x) && aligned16(y) && (ldx % 4 == 0) && (ldy % 4 == 0) &&
                   (!addend || (aligned16(addend) && ldadd % 4 == 0));
#define GTE_SPMM_GO(VEC, G, C) \
  return launch_spmm<VEC, G, C>(indptr, indices, w, pre_scale, row_norm, mode, x, ldx, addend, ldadd, y, ldy, n_rows, f, st)
  if (vec) {
    const int nv = (f + 3) / 4;  // 128-bit chunks per row (padding columns in [f, ld) are touched, never interpreted)
    if (nv <= 4) GTE_SPMM_GO(4, 4, 1);
    if (nv <= 8) GTE_SPMM_GO(4, 8, 1);
    if (nv <= 16) GTE_SPMM_GO(4, 16, 1);
    if (nv <= 32) GTE_SPMM_GO(4, 32, 1);
    if (nv <= 64) GTE_SPMM_GO(4, 32, 2);
    if (nv <= 96) GTE_SPMM_GO(4, 32, 3);
    GTE_SPMM_GO(4, 32, 4);  // wider rows tile over grid.y in blocks of 512 columns
  } else {
    if (f <= 4) GTE_SPMM_GO(1, 4, 1);
    if (f <= 8) GTE_SPMM_GO(1, 8, 1);
    if (f <= 16) GTE_SPMM_GO(1, 16, 1);
    if (f <= 32) GTE_SPMM_GO(1, 32, 1);
    if (f <= 64) GTE_SPMM_GO(1, 32, 2);
    GTE_SPMM_GO(1, 32, 4);  // blocks of 128 columns over grid.y
  }
#undef GTE_SPMM_GO
}
