// Dense projection of the GraphSAGE layers: nn.Linear over cat(h, ah*norm)
// (/root/reference/src/components/graphs/models.py:27,63,69-72) and its
// autograd (input gradient, weight/bias gradient), on the FFMA GEMM of
// gte_gemm.cuh.  The tensor-core route for wide hidden layers lives in
// gte_umma.cu and is selected by the host layer; this file is the exact-fp32
// path and serves every narrow shape.
#include "gte_gemm.cuh"
#include "gte_gram.cuh"

namespace gte {

// out(m,n) (+)= sum_z partial[z][m][n] in a fixed order (deterministic); 32 outputs per block
__global__ void __launch_bounds__(RED_THREADS)
    k_splitk_reduce(const float* __restrict__ partial, int splits, int32_t M, int32_t N, int64_t split_stride,
                    float* __restrict__ out, int64_t so_m, int64_t so_n, int accumulate) {
  __shared__ float red[RED_THREADS];
  const int64_t i = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
  const bool valid = i < (int64_t)M * N;
  float s = reduce_partials_block(partial, splits, split_stride, i, valid, red);
  if ((threadIdx.x >> 5) != 0 || !valid) return;
  const int32_t m = (int32_t)(i / N), n = (int32_t)(i % N);
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


// ---- one column block of the weight gradient: dWblk[o][j] (+)= sum_r dz[r,o] * x[r,j] -----------------------
static size_t dw_block_ws_bytes(int32_t n, int32_t fo, int32_t k, bool /*want_ones*/) {
  if (k == 0) return 0;
  // upper bound over every route dw_block may take for this shape
  const int32_t M = fo >= k ? fo : k, N = fo >= k ? k : fo;
  SplitPlan p = plan_split(n, M, N);
  size_t need = up256((size_t)p.splits * M * N * 4);
  auto upd = [&](size_t v) { if (v > need) need = v; };
  if (gram_eligible(k)) upd(gram_plan(n, fo, k).ws_bytes);
  if (gram_eligible(k + 1)) upd(gram_plan(n, fo, k + 1).ws_bytes);
  if (gram_eligible(fo)) upd(gram_plan(n, k, fo).ws_bytes);
  return need;
}

// returns GTE_OK; *did_ones is set when the bias gradient was produced by the same pass
static int dw_block(const float* dz, int64_t lddz, int32_t fo, const float* x, int64_t ldx, int32_t k, float* dWblk,
                    int64_t lddw, float* db, int accumulate, int32_t n, float* part, cudaStream_t st, bool* did_ones) {
  *did_ones = false;
  if (k == 0) return GTE_OK;
  if (gram_eligible(k + (db ? 1 : 0))) {  // narrow x (input layer): stream dz once, bias gradient for free
    *did_ones = db != nullptr;
    return gram_tall(dz, lddz, fo, x, ldx, k, nullptr, 0, 0, db != nullptr, n, dWblk, lddw, 1, nullptr, 0, 0, db, nullptr,
                     accumulate, part, st);
  }
  if (gram_eligible(fo))  // narrow dz (class layer): stream x once, write the transpose
    return gram_tall(x, ldx, k, dz, lddz, fo, nullptr, 0, 0, false, n, dWblk, 1, lddw, nullptr, 0, 0, nullptr, nullptr,
                     accumulate, part, st);
  const bool dz_is_A = fo >= k;  // put the wider side on the 128-row tile axis
  const int32_t M = dz_is_A ? fo : k, N = dz_is_A ? k : fo;
  SplitPlan p = plan_split(n, M, N);
  GemmArgs g{};
  const float* Aptr = dz_is_A ? dz : x;
  const int64_t lda = dz_is_A ? lddz : ldx;
  const float* Bptr = dz_is_A ? x : dz;
  const int64_t ldb = dz_is_A ? ldx : lddz;
  g.seg[0] = GemmSeg{Aptr, lda, Bptr, ldb, n, aligned16(Aptr) && lda % 4 == 0, aligned16(Bptr) && ldb % 4 == 0};
  g.nseg = 1;
  g.M = M;
  g.N = N;
  g.C = part;
  g.c_stride_m = N;
  g.c_stride_n = 1;
  g.vecC = aligned16(part) && N % 4 == 0;
  g.k_chunk = p.k_chunk;
  g.split_stride = (int64_t)M * N;
  int rc = launch_gemm<false, false>(g, p.splits, st);
  if (rc != GTE_OK) return rc;
  const int64_t so_m = dz_is_A ? lddw : 1;
  const int64_t so_n = dz_is_A ? 1 : lddw;
  const int64_t total = (int64_t)M * N;
  k_splitk_reduce<<<(unsigned)ceil_div64(total, 32), RED_THREADS, 0, st>>>(part, p.splits, M, N, (int64_t)M * N, dWblk, so_m, so_n,
                                                                   accumulate);
  GTE_CHECK_LAUNCH("k_splitk_reduce");
  return GTE_OK;
}

static size_t colsum_ws_bytes(int32_t n, int32_t fo) {
  SplitPlan pb = plan_split(n, 128, 128);
  return up256((size_t)pb.splits * (size_t)(fo > 0 ? fo : 1) * 4);
}

static int colsum(const float* dz, int64_t lddz, int32_t n, int32_t fo, float* db, int accumulate, float* bpart,
                  cudaStream_t st) {
  SplitPlan pb = plan_split(n, 128, 128);
  if (n > 0) {
    k_colsum_partial<<<pb.splits, 256, 0, st>>>(dz, lddz, n, fo, pb.k_chunk, bpart);
    GTE_CHECK_LAUNCH("k_colsum_partial");
  }
  k_splitk_reduce<<<(unsigned)ceil_div64(fo, 32), RED_THREADS, 0, st>>>(bpart, n > 0 ? pb.splits : 0, 1, fo, (int64_t)fo, db, 0, 1,
                                                                accumulate);
  GTE_CHECK_LAUNCH("k_splitk_reduce(db)");
  return GTE_OK;
}

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
  g.vecC = aligned16(z) && ldz % 4 == 0;
  g.bias = bias;
  return launch_gemm<true, true>(g, 0, as_stream(stream));
}

int gte_linear_bwd_data2(const float* dz1, int64_t lddz1, int32_t col1, const float* dz2, int64_t lddz2, int32_t col2,
                         int32_t fo, const float* W, int64_t ldw, int32_t k, const float* row_scale, float* dx,
                         int64_t lddx, int32_t n, int accumulate, gte_stream_t stream) {
  GTE_CHECK_ARG(n >= 0 && fo >= 0 && k >= 0 && col1 >= 0 && col2 >= 0, "gte_linear_bwd_data: negative size");
  if (n == 0 || k == 0) return GTE_OK;
  GTE_CHECK_ARG(dz1 && W && dx, "gte_linear_bwd_data: null argument");
  GTE_CHECK_ARG(lddz1 >= fo && (!dz2 || lddz2 >= fo) && lddx >= k && ldw >= (int64_t)col1 + k &&
                    (!dz2 || ldw >= (int64_t)col2 + k),
                "gte_linear_bwd_data: leading dimension too small");
  GemmArgs g{};
  const float* W1 = W + col1;
  g.seg[0] = GemmSeg{dz1, lddz1, W1, ldw, fo, aligned16(dz1) && lddz1 % 4 == 0, aligned16(W1) && ldw % 4 == 0};
  g.nseg = 1;
  if (dz2) {
    const float* W2 = W + col2;
    g.seg[1] = GemmSeg{dz2, lddz2, W2, ldw, fo, aligned16(dz2) && lddz2 % 4 == 0, aligned16(W2) && ldw % 4 == 0};
    g.nseg = 2;
  }
  g.M = n;
  g.N = k;
  g.C = dx;
  g.c_stride_m = lddx;
  g.c_stride_n = 1;
  g.vecC = aligned16(dx) && lddx % 4 == 0;
  g.row_scale = row_scale;
  g.accumulate = accumulate;
  return launch_gemm<true, false>(g, 0, as_stream(stream));
}

int gte_linear_bwd_data(const float* dz, int64_t lddz, int32_t fo, const float* W, int64_t ldw, int32_t col0,
                        int32_t k, const float* row_scale, float* dx, int64_t lddx, int32_t n, int accumulate,
                        gte_stream_t stream) {
  return gte_linear_bwd_data2(dz, lddz, col0, nullptr, 0, 0, fo, W, ldw, k, row_scale, dx, lddx, n, accumulate, stream);
}

size_t gte_linear_bwd_weight_workspace_bytes(int32_t n, int32_t fo, int32_t k1, int32_t k2) {
  if (n < 0 || fo < 0 || k1 < 0 || k2 < 0) return 0;
  size_t a = dw_block_ws_bytes(n, fo, k1, true), b = dw_block_ws_bytes(n, fo, k2, true);
  size_t g2 = gram_eligible(2 * fo) ? gram_plan(n, k1 > k2 ? k1 : k2, 2 * fo).ws_bytes : 0;
  size_t need = a > b ? a : b;
  if (g2 > need) need = g2;
  if (gram_eligible(k1 + k2 + 1)) {
    const size_t g3 = gram_plan(n, fo, k1 + k2 + 1).ws_bytes, g4 = gram_plan(n, fo, k1 + k2).ws_bytes;
    if (g3 > need) need = g3;
    if (g4 > need) need = g4;
  }
  return need + colsum_ws_bytes(n, fo) + 256;
}

int gte_linear_bwd_weight(const float* dz, int64_t lddz, int32_t fo, const float* x1, int64_t ldx1, int32_t k1,
                          const float* x2, int64_t ldx2, int32_t k2, float* dW, int64_t lddw, float* db,
                          int accumulate, int32_t n, void* ws, size_t ws_bytes, gte_stream_t stream) {
  GTE_CHECK_ARG(n >= 0 && fo >= 0 && k1 >= 0 && k2 >= 0, "gte_linear_bwd_weight: negative size");
  if (fo == 0) return GTE_OK;
  GTE_CHECK_ARG(dW != nullptr, "gte_linear_bwd_weight: dW is null");
  GTE_CHECK_ARG(n == 0 || dz, "gte_linear_bwd_weight: dz is null");
  GTE_CHECK_ARG(lddw >= (int64_t)k1 + k2 && lddz >= fo, "gte_linear_bwd_weight: leading dimension too small");
  GTE_CHECK_ARG(n == 0 || k1 == 0 || (x1 && ldx1 >= k1), "gte_linear_bwd_weight: bad x1");
  GTE_CHECK_ARG(n == 0 || k2 == 0 || (x2 && ldx2 >= k2), "gte_linear_bwd_weight: bad x2");
  const size_t need = gte_linear_bwd_weight_workspace_bytes(n, fo, k1, k2);
  if (ws == nullptr || ws_bytes < need)
    return fail(GTE_ERR_WORKSPACE, "gte_linear_bwd_weight: workspace %zu < required %zu", ws_bytes, need);
  cudaStream_t st = as_stream(stream);
  float* part = static_cast<float*>(ws);
  if (k1 > 0 && k2 > 0 && gram_eligible(k1 + k2 + (db ? 1 : 0)))  // both blocks narrow (input layer): ONE pass over dz
    return gram_tall(dz, lddz, fo, x1, ldx1, k1, x2, ldx2, k2, db != nullptr, n, dW, lddw, 1, dW + k1, lddw, 1, db, nullptr,
                     accumulate, part, st);
  bool have_db = db == nullptr, did = false;
  int rc = dw_block(dz, lddz, fo, x1, ldx1, k1, dW, lddw, have_db ? nullptr : db, accumulate, n, part, st, &did);
  if (rc != GTE_OK) return rc;
  have_db = have_db || did;
  rc = dw_block(dz, lddz, fo, x2, ldx2, k2, dW + k1, lddw, have_db ? nullptr : db, accumulate, n, part, st, &did);
  if (rc != GTE_OK) return rc;
  have_db = have_db || did;
  if (!have_db) {
    float* bpart = reinterpret_cast<float*>(static_cast<char*>(ws) + (need - 256 - colsum_ws_bytes(n, fo)));
    return colsum(dz, lddz, n, fo, db, accumulate, bpart, st);
  }
  return GTE_OK;
}

int gte_linear_bwd_weight2(const float* dz1, int64_t lddz1, const float* dz2, int64_t lddz2, int32_t fo, const float* x,
                           int64_t ldx, int32_t k, float* dW, int64_t lddw, int32_t col1, int32_t col2, float* db,
                           int accumulate, int32_t n, void* ws, size_t ws_bytes, gte_stream_t stream) {
  GTE_CHECK_ARG(n >= 0 && fo >= 0 && k >= 0 && col1 >= 0 && col2 >= 0, "gte_linear_bwd_weight2: negative size");
  if (fo == 0 || k == 0) return GTE_OK;
  GTE_CHECK_ARG(dW != nullptr, "gte_linear_bwd_weight2: dW is null");
  GTE_CHECK_ARG(n == 0 || (dz1 && dz2 && x), "gte_linear_bwd_weight2: null argument");
  GTE_CHECK_ARG(lddz1 >= fo && lddz2 >= fo && ldx >= k && lddw >= (int64_t)col1 + k && lddw >= (int64_t)col2 + k,
                "gte_linear_bwd_weight2: leading dimension too small");
  const size_t need = gte_linear_bwd_weight_workspace_bytes(n, fo, k, k);
  if (ws == nullptr || ws_bytes < need)
    return fail(GTE_ERR_WORKSPACE, "gte_linear_bwd_weight2: workspace %zu < required %zu", ws_bytes, need);
  cudaStream_t st = as_stream(stream);
  float* part = static_cast<float*>(ws);
  float* bpart = reinterpret_cast<float*>(static_cast<char*>(ws) + (need - 256 - colsum_ws_bytes(n, fo)));
  if (gram_eligible(2 * fo)) {
    // one pass over the wide operand x for both gradient blocks: Q = [dz1 | dz2]
    const bool fuse_db = db && aligned16(x) && ldx % 4 == 0;  // bias gradient = column sums of dz1, from the same pass
    int rc = gram_tall(x, ldx, k, dz1, lddz1, fo, dz2, lddz2, fo, false, n, dW + col1, 1, lddw, dW + col2, 1, lddw, nullptr,
                       fuse_db ? db : nullptr, accumulate, part, st);
    if (rc != GTE_OK) return rc;
    return (db && !fuse_db) ? colsum(dz1, lddz1, n, fo, db, accumulate, bpart, st) : GTE_OK;
  }
  bool did = false;
  int rc = dw_block(dz1, lddz1, fo, x, ldx, k, dW + col1, lddw, nullptr, accumulate, n, part, st, &did);
  if (rc != GTE_OK) return rc;
  rc = dw_block(dz2, lddz2, fo, x, ldx, k, dW + col2, lddw, nullptr, accumulate, n, part, st, &did);
  if (rc != GTE_OK) return rc;
  return db ? colsum(dz1, lddz1, n, fo, db, accumulate, bpart, st) : GTE_OK;
}

}  // extern "C"
