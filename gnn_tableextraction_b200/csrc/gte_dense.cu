// Dense projection of the GraphSAGE layers: nn.Linear over cat(h, ah*norm)
// (/root/reference/src/components/graphs/models.py:27,63,69-72) and its
// autograd (input gradient, weight/bias gradient), on the FFMA GEMM of
// gte_gemm.cuh.  The tensor-core route for wide hidden layers lives in
// gte_umma.cu and is selected by the host layer; this file is the exact-fp32
// path and serves every narrow shape.
#include "gte_gemm.cuh"

namespace gte {

// out(m,n) (+)= sum_z partial[z][m][n], z ascending (fixed order => deterministic)
__global__ void k_splitk_reduce(const float* __restrict__ partial, int splits, int32_t M, int32_t N,
                                int64_t split_stride, float* __restrict__ out, int64_t so_m, int64_t so_n,
                                int accumulate) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  const int32_t m = (int32_t)(i / N), n = (int32_t)(i % N);
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += partial[(int64_t)z * split_stride + i];
  float* p = out + m * so_m + n * so_n;
  if (accumulate) s += *p;
  *p = s;
}

// partial[b][c] = sum over rows [b*rows_per_block, ...) of dz[r][c]
__global__ void k_colsum_partial(const float* __restrict__ dz, int64_t ld, int32_t n, int32_t f,
                                 int32_t rows_per_block, float* __restrict__ partial) {
  const int32_t r0 = blockIdx.x * rows_per_block;
  const int32_t r1 = min(n, r0 + rows_per_block);
  for (int32_t c = threadIdx.x; c < f; c += blockDim.x) {
    float s = 0.f;
    for (int32_t r = r0; r < r1; ++r) s += __ldg(dz + (int64_t)r * ld + c);
    partial[(int64_t)blockIdx.x * f + c] = s;
  }
}

struct SplitPlan {
  int splits;
  int32_t k_chunk;
};

// Depends only on the problem shape (not on the device) so results reproduce.
static SplitPlan plan_split(int32_t rows, int32_t M, int32_t N) {
  const int64_t tiles = ceil_div64(M, 128) * ceil_div64(N, N <= 16 ? 16 : (N <= 32 ? 32 : (N <= 64 ? 64 : 128)));
  int64_t want = ceil_div64(592, tiles);  // ~4 CTAs per SM on a 148-SM part
  if (want > 256) want = 256;
  int64_t chunk = ceil_div64(rows, want);
  if (chunk < 256) chunk = 256;
  chunk = ceil_div64(chunk, GEMM_BK) * GEMM_BK;
  SplitPlan p;
  p.k_chunk = (int32_t)chunk;
  p.splits = (int)ceil_div64(rows, chunk);
  if (p.splits < 1) p.splits = 1;
  return p;
}

static size_t up256(size_t v) { return (v + 255) & ~size_t(255); }

}  // namespace gte

using namespace gte;

extern "C" {

int gte_linear_fwd(const float* x1, int64_t ldx1, int32_t k1, const float* x2, int64_t ldx2, int32_t k2,
                   const float* W, int64_t ldw, const float* bias, float* z, int64_t ldz, int32_t n, int32_t fo,
                   gte_stream_t stream) {
  GTE_CHECK_ARG(n >= 0 && fo >= 0 && k1 >= 0 && k2 >= 0, "gte_linear_fwd: negative size");
  if (n == 0 || fo == 0) return GTE_OK;
  GTE_CHECK_ARG(W && z, "gte_linear_fwd: null W or z");
  GTE_CHECK_ARG(k1 == 0 || (x1 && ldx1 >= k1), "gte_linear_fwd: bad x1");
  GTE_CHECK_ARG(k2 == 0 || (x2 && ldx2 >= k2), "gte_linear_fwd: bad x2");
  GTE_CHECK_ARG(ldw >= (int64_t)k1 + k2 && ldz >= fo, "gte_linear_fwd: leading dimension too small");
  GemmArgs g{};
  int s = 0;
  if (k1 > 0) {
    g.seg[s] = GemmSeg{x1, ldx1, W, ldw, k1, aligned16(x1) && ldx1 % 4 == 0, aligned16(W) && ldw % 4 == 0};
    ++s;
  }
  if (k2 > 0) {
    g.seg[s] = GemmSeg{x2, ldx2, W + k1, ldw, k2, aligned16(x2) && ldx2 % 4 == 0, aligned16(W + k1) && ldw % 4 == 0};
    ++s;
  }
  g.nseg = s;
  g.M = n;
  g.N = fo;
  g.C = z;
  g.c_stride_m = ldz;
  g.c_stride_n = 1;
  g.bias = bias;
  return launch_gemm<true, true>(g, 0, as_stream(stream));
}

int gte_linear_bwd_data(const float* dz, int64_t lddz, int32_t fo, const float* W, int64_t ldw, int32_t col0,
                        int32_t k, const float* row_scale, float* dx, int64_t lddx, int32_t n, int accumulate,
                        gte_stream_t stream) {
  GTE_CHECK_ARG(n >= 0 && fo >= 0 && k >= 0 && col0 >= 0, "gte_linear_bwd_data: negative size");
  if (n == 0 || k == 0) return GTE_OK;
  GTE_CHECK_ARG(dz && W && dx, "gte_linear_bwd_data: null argument");
  GTE_CHECK_ARG(lddz >= fo && lddx >= k && ldw >= (int64_t)col0 + k, "gte_linear_bwd_data: leading dimension too small");
  GemmArgs g{};
  const float* Wb = W + col0;
  g.seg[0] = GemmSeg{dz, lddz, Wb, ldw, fo, aligned16(dz) && lddz % 4 == 0, aligned16(Wb) && ldw % 4 == 0};
  g.nseg = 1;
  g.M = n;
  g.N = k;
  g.C = dx;
  g.c_stride_m = lddx;
  g.c_stride_n = 1;
  g.row_scale = row_scale;
  g.accumulate = accumulate;
  return launch_gemm<true, false>(g, 0, as_stream(stream));
}

size_t gte_linear_bwd_weight_workspace_bytes(int32_t n, int32_t fo, int32_t k1, int32_t k2) {
  if (n < 0 || fo < 0 || k1 < 0 || k2 < 0) return 0;
  size_t need = 0;
  const int32_t ks[2] = {k1, k2};
  for (int s = 0; s < 2; ++s) {
    if (ks[s] == 0) continue;
    const int32_t M = fo >= ks[s] ? fo : ks[s], N = fo >= ks[s] ? ks[s] : fo;
    SplitPlan p = plan_split(n, M, N);
    size_t b = up256((size_t)p.splits * M * N * 4);
    if (b > need) need = b;
  }
  SplitPlan pb = plan_split(n, 128, 128);
  need += up256((size_t)pb.splits * (size_t)(fo > 0 ? fo : 1) * 4);
  return need + 256;
}

int gte_linear_bwd_weight(const float* dz, int64_t lddz, int32_t fo, const float* x1, int64_t ldx1, int32_t k1,
                          const float* x2, int64_t ldx2, int32_t k2, float* dW, int64_t lddw, float* db,
                          int accumulate, int32_t n, void* ws, size_t ws_bytes, gte_stream_t stream) {
  GTE_CHECK_ARG(n >= 0 && fo >= 0 && k1 >= 0 && k2 >= 0, "gte_linear_bwd_weight: negative size");
  if (fo == 0) return GTE_OK;
  GTE_CHECK_ARG(dW != nullptr, "gte_linear_bwd_weight: dW is null");
  GTE_CHECK_ARG(n == 0 || dz, "gte_linear_bwd_weight: dz is null");
  GTE_CHECK_ARG(lddw >= (int64_t)k1 + k2 && lddz >= fo, "gte_linear_bwd_weight: leading dimension too small");
  const size_t need = gte_linear_bwd_weight_workspace_bytes(n, fo, k1, k2);
  if (ws == nullptr || ws_bytes < need)
    return fail(GTE_ERR_WORKSPACE, "gte_linear_bwd_weight: workspace %zu < required %zu", ws_bytes, need);
  cudaStream_t st = as_stream(stream);
  float* part = static_cast<float*>(ws);

  const float* xs[2] = {x1, x2};
  const int64_t ldxs[2] = {ldx1, ldx2};
  const int32_t ks[2] = {k1, k2};
  size_t gemm_region = 0;
  int32_t col_off = 0;
  for (int s = 0; s < 2; ++s) {
    const int32_t k = ks[s];
    if (k == 0) continue;
    GTE_CHECK_ARG(n == 0 || (xs[s] && ldxs[s] >= k), "gte_linear_bwd_weight: bad x%d", s + 1);
    const bool dz_is_A = fo >= k;  // put the wider side on the 128-row tile axis
    const int32_t M = dz_is_A ? fo : k, N = dz_is_A ? k : fo;
    SplitPlan p = plan_split(n, M, N);
    size_t b = up256((size_t)p.splits * M * N * 4);
    if (b > gemm_region) gemm_region = b;
    GemmArgs g{};
    const float* Aptr = dz_is_A ? dz : xs[s];
    const int64_t lda = dz_is_A ? lddz : ldxs[s];
    const float* Bptr = dz_is_A ? xs[s] : dz;
    const int64_t ldb = dz_is_A ? ldxs[s] : lddz;
    g.seg[0] = GemmSeg{Aptr, lda, Bptr, ldb, n, aligned16(Aptr) && lda % 4 == 0, aligned16(Bptr) && ldb % 4 == 0};
    g.nseg = 1;
    g.M = M;
    g.N = N;
    g.C = part;
    g.c_stride_m = N;
    g.c_stride_n = 1;
    g.k_chunk = p.k_chunk;
    g.split_stride = (int64_t)M * N;
    int rc = launch_gemm<false, false>(g, p.splits, st);
    if (rc != GTE_OK) return rc;
    // partial(m,n) -> dW[o][col_off + j]
    const int64_t so_m = dz_is_A ? lddw : 1;
    const int64_t so_n = dz_is_A ? 1 : lddw;
    const int64_t total = (int64_t)M * N;
    k_splitk_reduce<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(part, p.splits, M, N, (int64_t)M * N,
                                                                     dW + col_off, so_m, so_n, accumulate);
    GTE_CHECK_LAUNCH("k_splitk_reduce");
    col_off += k;
  }
  if (db) {
    // recompute the region size exactly as the workspace query does
    size_t region = 0;
    for (int s = 0; s < 2; ++s) {
      if (ks[s] == 0) continue;
      const int32_t M = fo >= ks[s] ? fo : ks[s], N = fo >= ks[s] ? ks[s] : fo;
      SplitPlan p = plan_split(n, M, N);
      size_t b = up256((size_t)p.splits * M * N * 4);
      if (b > region) region = b;
    }
    float* bpart = reinterpret_cast<float*>(static_cast<char*>(ws) + region);
    SplitPlan pb = plan_split(n, 128, 128);
    if (n > 0) {
      k_colsum_partial<<<pb.splits, 256, 0, st>>>(dz, lddz, n, fo, pb.k_chunk, bpart);
      GTE_CHECK_LAUNCH("k_colsum_partial");
    }
    k_splitk_reduce<<<(unsigned)ceil_div64(fo, 256), 256, 0, st>>>(bpart, n > 0 ? pb.splits : 0, 1, fo, (int64_t)fo, db,
                                                                  0, 1, accumulate);
    GTE_CHECK_LAUNCH("k_splitk_reduce(db)");
  }
  return GTE_OK;
}

}  // extern "C"
