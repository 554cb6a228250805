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

// ----------------------------------------------------------------------------------------------
// Block-diagonal (page-batched) variant: one CTA per (page, column slice).  A batch of page graphs has
// no edge between pages, so every source row a page's rows can touch lies in that page's own node
// range.  The CTA stages, once and coalesced, everything the page needs in shared memory:
//   x[page rows, slice] (cp.async), the page's row pointers, its column indices (made page-local) and
//   its edge weights (with the source-side scale folded in),
// after which the whole aggregation of the page runs out of shared memory: no dependent global
// load is left in the per-row loop, and the L2->SM gather traffic (4*E*F bytes, ~10x the compulsory
// bytes at degree 10) disappears -- the kernel streams x in / y out at HBM rate.  Indices outside
// the page window (not block diagonal) fall back to a global load, so the result is always correct.
constexpr int PAGED_THREADS = 512;

template <int G>
__global__ void __launch_bounds__(PAGED_THREADS)
    k_spmm_paged(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ w,
                 const float* __restrict__ pre_scale, const float* __restrict__ row_norm, int mode,
                 const float* __restrict__ x, int64_t ldx, const float* __restrict__ addend, int64_t ldadd,
                 float* __restrict__ y, int64_t ldy, const int32_t* __restrict__ page_off, int32_t f,
                 int32_t np_cap, int32_t ne_cap) {
  extern __shared__ __align__(16) float smem_f[];
  constexpr int CS = G * 4;
  constexpr int ROWS_PER_ITER = PAGED_THREADS / G;
  float* sx = smem_f;                                                     // [np_cap][CS]
  uint32_t* s_off = reinterpret_cast<uint32_t*>(sx + (size_t)np_cap * CS);  // [ne_cap] byte offset of the source row in sx
  float* s_w = reinterpret_cast<float*>(s_off + ne_cap);                  // [ne_cap] edge weight (* source-side scale)
  int32_t* s_ptr = reinterpret_cast<int32_t*>(s_w + ne_cap);              // [np_cap + 1] page-local row pointers
  int32_t* s_flag = s_ptr + np_cap + 1;                                   // some edge needs the global slow path
  // the slices of one page are adjacent in launch order (blockIdx.x fastest) so that cache lines shared
  // by neighbouring slices are fetched from DRAM once
  const int page = blockIdx.y;
  const int c0 = blockIdx.x * CS;
  const int32_t n0 = page_off[page], n1 = page_off[page + 1];
  const int32_t np = n1 - n0;
  const int32_t e0 = indptr[n0];
  const int32_t ne = indptr[n1] - e0;
  const bool staged_edges = ne <= ne_cap;  // always true when the caller's max_page_edges is right
  if (threadIdx.x == 0) *s_flag = staged_edges ? 0 : 1;
  for (int i = threadIdx.x; i < np * G; i += PAGED_THREADS) {
    const int r = i / G, ch = i % G;
    const int col = c0 + ch * 4;
    float* dst = sx + (size_t)r * CS + ch * 4;
    if (col < f) {  // 16-byte chunks; chunks entirely beyond f are zero-filled
      const float* src = x + (int64_t)(n0 + r) * ldx + col;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                   : "memory");
    } else {
      *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = threadIdx.x; i <= np; i += PAGED_THREADS) s_ptr[i] = indptr[n0 + i] - e0;
  __syncthreads();  // s_flag initialised before anyone raises it
  if (staged_edges) {
    for (int i = threadIdx.x; i < ne; i += PAGED_THREADS) {
      const int32_t c = __ldg(indices + e0 + i);
      float wv = w ? __ldg(w + e0 + i) : 1.0f;
      if (pre_scale) wv *= __ldg(pre_scale + c);
      const int32_t sl = c - n0;
      const bool inside = sl >= 0 && sl < np;
      s_off[i] = inside ? (uint32_t)sl * (CS * 4) : 0u;
      s_w[i] = inside ? wv : 0.f;  // outside edges contribute through the slow path below
      if (!inside) *s_flag = 1;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int lane = threadIdx.x % G;
  const int grp = threadIdx.x / G;
  const int col = c0 + lane * 4;
  const bool on = col < f;
  const bool slow = *s_flag != 0;
  const char* xb = reinterpret_cast<const char*>(sx) + lane * 16;
  for (int32_t rl = grp; rl < np; rl += ROWS_PER_ITER) {
    const int64_t row = n0 + rl;
    const int32_t beg = s_ptr[rl], end = s_ptr[rl + 1];
    // issue the row-end operands early so their latency overlaps the neighbour loop
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
    if (addend && on) av = __ldg(reinterpret_cast<const float4*>(addend + row * ldadd + col));
    const float nrm = (mode == GTE_AGG_SUM_NORM) ? __ldg(row_norm + row) : 1.0f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (staged_edges) {
      int32_t j = beg;
      for (; j + 4 <= end; j += 4) {  // edge order preserved: deterministic, same sums as the generic kernel
        const uint32_t o0 = s_off[j], o1 = s_off[j + 1], o2 = s_off[j + 2], o3 = s_off[j + 3];
        const float w0 = s_w[j], w1 = s_w[j + 1], w2 = s_w[j + 2], w3 = s_w[j + 3];
        const float4 x0 = *reinterpret_cast<const float4*>(xb + o0);
        const float4 x1 = *reinterpret_cast<const float4*>(xb + o1);
        const float4 x2 = *reinterpret_cast<const float4*>(xb + o2);
        const float4 x3 = *reinterpret_cast<const float4*>(xb + o3);
        acc.x = fmaf(w0, x0.x, acc.x); acc.y = fmaf(w0, x0.y, acc.y); acc.z = fmaf(w0, x0.z, acc.z); acc.w = fmaf(w0, x0.w, acc.w);
        acc.x = fmaf(w1, x1.x, acc.x); acc.y = fmaf(w1, x1.y, acc.y); acc.z = fmaf(w1, x1.z, acc.z); acc.w = fmaf(w1, x1.w, acc.w);
        acc.x = fmaf(w2, x2.x, acc.x); acc.y = fmaf(w2, x2.y, acc.y); acc.z = fmaf(w2, x2.z, acc.z); acc.w = fmaf(w2, x2.w, acc.w);
        acc.x = fmaf(w3, x3.x, acc.x); acc.y = fmaf(w3, x3.y, acc.y); acc.z = fmaf(w3, x3.z, acc.z); acc.w = fmaf(w3, x3.w, acc.w);
      }
      for (; j < end; ++j) {
        const float wt = s_w[j];
        const float4 xv = *reinterpret_cast<const float4*>(xb + s_off[j]);
        acc.x = fmaf(wt, xv.x, acc.x); acc.y = fmaf(wt, xv.y, acc.y); acc.z = fmaf(wt, xv.z, acc.z); acc.w = fmaf(wt, xv.w, acc.w);
      }
    }
    if (slow && on) {  // rare: the page table does not describe a block-diagonal graph (or edge capacity too small)
      for (int32_t j = beg; j < end; ++j) {
        const int32_t c = __ldg(indices + e0 + j);
        const int32_t sl = c - n0;
        if (staged_edges && sl >= 0 && sl < np) continue;  // already accumulated from shared memory
        float wt = w ? __ldg(w + e0 + j) : 1.0f;
        if (pre_scale) wt *= __ldg(pre_scale + c);
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (int64_t)c * ldx + col));
        acc.x = fmaf(wt, xv.x, acc.x); acc.y = fmaf(wt, xv.y, acc.y); acc.z = fmaf(wt, xv.z, acc.z); acc.w = fmaf(wt, xv.w, acc.w);
      }
    }
    if (!on) continue;
    if (mode == GTE_AGG_MEAN) {
      const int32_t deg = end - beg;
      const float d = (float)(deg > 1 ? deg : 1);
      acc.x /= d; acc.y /= d; acc.z /= d; acc.w /= d;
    } else if (mode == GTE_AGG_SUM_NORM) {
      acc.x *= nrm; acc.y *= nrm; acc.z *= nrm; acc.w *= nrm;
    }
    if (addend) {
      acc.x += av.x; acc.y += av.y; acc.z += av.z; acc.w += av.w;
    }
    *reinterpret_cast<float4*>(y + row * ldy + col) = acc;
  }
}

static size_t paged_smem_bytes(int G, int32_t np_cap, int32_t ne_cap) {
  return (size_t)np_cap * G * 16 + (size_t)ne_cap * 8 + ((size_t)np_cap + 2) * 4 + 16;
}

template <int G>
static int launch_spmm_paged(const int32_t* indptr, const int32_t* indices, const float* w, const float* pre_scale,
                             const float* row_norm, int mode, const float* x, int64_t ldx, const float* addend,
                             int64_t ldadd, float* y, int64_t ldy, const int32_t* page_off, int32_t num_pages,
                             int32_t np_cap, int32_t ne_cap, int32_t f, cudaStream_t st) {
  constexpr int CS = G * 4;
  const size_t smem = paged_smem_bytes(G, np_cap, ne_cap);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    GTE_CHECK_CUDA(cudaFuncSetAttribute(k_spmm_paged<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                   "k_spmm_paged(smem attr)");
    configured = smem;
  }
  dim3 grid((unsigned)ceil_div64(f, CS), (unsigned)num_pages);
  k_spmm_paged<G><<<grid, PAGED_THREADS, smem, st>>>(indptr, indices, w, pre_scale, row_norm, mode, x, ldx, addend, ldadd, y,
                                                    ldy, page_off, f, np_cap, ne_cap);
  GTE_CHECK_LAUNCH("k_spmm_paged");
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

extern "C" int gte_spmm_paged(const int32_t* indptr, const int32_t* indices, const float* w, const float* pre_scale,
                              const float* row_norm, int mode, const float* x, int64_t ldx, const float* addend,
                              int64_t ldadd, float* y, int64_t ldy, const int32_t* page_off, int32_t num_pages,
                              int32_t max_page_nodes, int32_t max_page_edges, int32_t n_rows, int32_t f,
                              gte_stream_t stream) {
  GTE_CHECK_ARG(n_rows >= 0 && f >= 0 && num_pages >= 0 && max_page_nodes >= 0 && max_page_edges >= 0,
                "gte_spmm_paged: negative size");
  GTE_CHECK_ARG(mode == GTE_AGG_SUM || mode == GTE_AGG_SUM_NORM || mode == GTE_AGG_MEAN, "gte_spmm_paged: bad mode %d", mode);
  if (n_rows == 0 || f == 0 || num_pages == 0) return GTE_OK;
  GTE_CHECK_ARG(indptr && indices && x && y && page_off, "gte_spmm_paged: null argument");
  GTE_CHECK_ARG(mode != GTE_AGG_SUM_NORM || row_norm, "gte_spmm_paged: SUM_NORM needs row_norm");
  GTE_CHECK_ARG(ldx >= f && ldy >= f && (!addend || ldadd >= f), "gte_spmm_paged: leading dimension < f");
  GTE_CHECK_ARG(x != y, "gte_spmm_paged: x and y must not alias");
  const bool vec = aligned16(x) && aligned16(y) && (ldx % 4 == 0) && (ldy % 4 == 0) &&
                   (!addend || (aligned16(addend) && ldadd % 4 == 0));
  // widest column slice whose page working set fits ~110 KB of shared memory (2 CTAs per SM)
  const size_t budget = 110 * 1024;
  int G = 0;
  const int fv = (f + 3) / 4;  // 16-byte chunks per row
  for (int g : {32, 16, 8}) {
    if (paged_smem_bytes(g, max_page_nodes, max_page_edges) <= budget) {
      G = g;
      break;
    }
  }
  while (G > 8 && G / 2 >= fv) G /= 2;  // narrow rows: do not stage padding
  if (!vec || G == 0)  // unaligned operands or pages too large to stage: generic L2-gather kernel
    return gte_spmm(indptr, indices, w, pre_scale, row_norm, mode, x, ldx, addend, ldadd, y, ldy, n_rows, f, stream);
  cudaStream_t st = as_stream(stream);
#define GTE_PAGED_GO(GG)                                                                                               \
  return launch_spmm_paged<GG>(indptr, indices, w, pre_scale, row_norm, mode, x, ldx, addend, ldadd, y, ldy, page_off, \
                               num_pages, max_page_nodes, max_page_edges, f, st)
  if (G == 32) GTE_PAGED_GO(32);
  if (G == 16) GTE_PAGED_GO(16);
  GTE_PAGED_GO(8);
#undef GTE_PAGED_GO
}
