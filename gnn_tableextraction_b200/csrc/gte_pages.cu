// The two steps either side of the layers that SURVEY.md section 8(f) ranks next:
//
//   k_page_predictions : `preds = logits.argmax(dim=1)` and the per-page accuracy
//                        `sum(preds == labels) / g.num_nodes()` of the predict loop
//                        (/root/reference/src/models/model_predict.py:144-154), for a whole batch of pages at once:
//                        int32 predictions (first maximal index, like torch.argmax) and per-page correct counts.
//   k_bbox_features    : the 13 BBOX node features of /root/reference/src/components/nlp/bbox.py:49-54,57-111:
//                        9 shape values from the integer box [x0, y0, x1, y1] and the 4-bin character-class
//                        histogram from per-box (letters, digits, others) counts -- float64 arithmetic like the
//                        Python original, cast to float32 at the end (model_train.py:295 `.float()`).
//
// Both are integer / tiny-arithmetic streams; results are bit-exact against the reference semantics.
#include "gte_common.cuh"

namespace gte {

__device__ __forceinline__ int64_t page_label(const void* labels, int dtype, int64_t i) {
  if (dtype == GTE_LABEL_I64) return static_cast<const int64_t*>(labels)[i];
  if (dtype == GTE_LABEL_I32) return static_cast<const int32_t*>(labels)[i];
  return (int64_t)static_cast<const float*>(labels)[i];  // float32 labels, `.long()` truncation (model_train.py:327)
}

// one CTA per page: predictions of its nodes + the page's number of correct predictions
__global__ void __launch_bounds__(256)
    k_page_predictions(const float* __restrict__ logits, int64_t ld, int32_t c, const void* __restrict__ labels,
                       int label_dtype, const int32_t* __restrict__ page_off, int32_t* __restrict__ preds,
                       int32_t* __restrict__ page_correct) {
  __shared__ int red[8];
  const int page = blockIdx.x;
  const int32_t n0 = page_off[page], n1 = page_off[page + 1];
  int correct = 0;
  for (int32_t i = n0 + threadIdx.x; i < n1; i += blockDim.x) {
    const float* lr = logits + (int64_t)i * ld;
    float mx = lr[0];
    int arg = 0;
    bool nan = mx != mx;  // torch.argmax treats NaN as the maximum (first NaN wins)
    for (int j = 1; j < c && !nan; ++j) {
      const float v = lr[j];
      if (v != v) {
        arg = j;
        nan = true;
      } else if (v > mx) {
        mx = v;
        arg = j;
      }
    }
    preds[i] = arg;
    if (labels && page_label(labels, label_dtype, i) == arg) ++correct;
  }
  if (page_correct) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) correct += __shfl_xor_sync(0xffffffffu, correct, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = correct;
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
      page_correct[page] = s;
    }
  }
}

// bbox.py:49-54 get_shape + bbox.py:57-111 get_histogram, one thread per text box
__global__ void k_bbox_features(const int32_t* __restrict__ boxes /*[n,4]*/, const int32_t* __restrict__ counts /*[n,3]*/,
                                float* __restrict__ out, int64_t ldo, int32_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t x0 = boxes[4 * i], y0 = boxes[4 * i + 1], x1 = boxes[4 * i + 2], y1 = boxes[4 * i + 3];
  const int32_t w = x1 - x0, h = y1 - y0;
  // int(width/2): Python true division, then truncation toward zero
  const int32_t hw = (int32_t)((double)w / 2.0), hh = (int32_t)((double)h / 2.0);
  float* o = out + i * ldo;
  o[0] = (float)w;
  o[1] = (float)h;
  o[2] = (float)(x1 - hw);
  o[3] = (float)(y1 - hh);
  o[4] = (float)((double)w * (double)h);
  o[5] = (float)x0;
  o[6] = (float)y0;
  o[7] = (float)x1;
  o[8] = (float)y1;
  double hist[4] = {0.0, 0.0, 0.0, 0.0};
  const int32_t lit = counts[3 * i], num = counts[3 * i + 1], oth = counts[3 * i + 2];
  const int32_t tot = lit + num + oth;
  if (tot != 0) {
    hist[0] = (double)lit / (double)tot;
    hist[1] = (double)num / (double)tot;
    hist[2] = (double)oth / (double)tot;
    // "keep sum 1 after truncate": Python sum() adds left to right starting from 0 (bbox.py:100-103)
    const double s = ((0.0 + hist[0]) + hist[1]) + hist[2] + hist[3];
    if (s != 1.0) {
      const double diff = 1.0 - s;
      double mxv = hist[0];
      int mi = 0;  // list.index(max(...)): first maximal entry
      for (int k = 1; k < 4; ++k)
        if (hist[k] > mxv) {
          mxv = hist[k];
          mi = k;
        }
      hist[mi] = mxv + diff;
    }
  }
  if (hist[0] == 0.0 && hist[1] == 0.0 && hist[2] == 0.0) hist[3] = 1.0;
  o[9] = (float)hist[0];
  o[10] = (float)hist[1];
  o[11] = (float)hist[2];
  o[12] = (float)hist[3];
}

}  // namespace gte

using namespace gte;

extern "C" int gte_page_predictions(const float* logits, int64_t ld, int32_t n, int32_t c, const void* labels,
                                    int label_dtype, const int32_t* page_off, int32_t num_pages, int32_t* preds,
                                    int32_t* page_correct, gte_stream_t stream) {
  GTE_CHECK_ARG(n >= 0 && c > 0 && num_pages >= 0, "gte_page_predictions: bad size");
  GTE_CHECK_ARG(label_dtype >= GTE_LABEL_I64 && label_dtype <= GTE_LABEL_F32, "gte_page_predictions: bad label dtype");
  if (n == 0 || num_pages == 0) return GTE_OK;
  GTE_CHECK_ARG(logits && page_off && preds && ld >= c, "gte_page_predictions: bad argument");
  GTE_CHECK_ARG(!page_correct || labels, "gte_page_predictions: page_correct needs labels");
  k_page_predictions<<<num_pages, 256, 0, as_stream(stream)>>>(logits, ld, c, labels, label_dtype, page_off, preds,
                                                              page_correct);
  GTE_CHECK_LAUNCH("k_page_predictions");
  return GTE_OK;
}

extern "C" int gte_bbox_features(const int32_t* boxes, const int32_t* counts, int32_t n, float* out, int64_t ldo,
                                 gte_stream_t stream) {
  GTE_CHECK_ARG(n >= 0, "gte_bbox_features: negative size");
  if (n == 0) return GTE_OK;
  GTE_CHECK_ARG(boxes && counts && out && ldo >= 13, "gte_bbox_features: bad argument");
  k_bbox_features<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(boxes, counts, out, ldo, n);
  GTE_CHECK_LAUNCH("k_bbox_features");
  return GTE_OK;
}

// ---------------------------------------------------------------------------------------------------
// Batch assembly of a page batch in ONE kernel (replaces `dgl.batch(...)`'s lazy
// COO -> CSC and COO -> CSR builds of /root/reference/src/models/model_train.py:297 + the per-layer
// `in_degrees` / `get_norm` of models.py:74-78):  one CTA per page sorts the page's edges by destination (CSC,
// forward) and by source (CSR, backward) entirely in shared memory -- stable, i.e. bit-identical to
// gte_csx_from_coo / a stable argsort of the batched COO -- and emits, in the same pass, the degree normaliser
// and the packed edge entries the conv-layer kernel stages (gte_paged_pack_edges' output for both directions).
// dgl.batch's contract is assumed: nodes and edges of page p occupy [page_off[p], page_off[p+1]) and
// [edge_off[p], edge_off[p+1]); an edge with an endpoint outside its page raises *bad (results undefined).
namespace gte {

constexpr int PF_THREADS = 256;

struct PfOut {
  int32_t *indptr, *indices, *eid;
  uint2* packed;
};

__global__ void __launch_bounds__(PF_THREADS)
    k_build_page_formats(const int32_t* __restrict__ src, const int32_t* __restrict__ dst, const float* __restrict__ w,
                         const int32_t* __restrict__ page_off, const int32_t* __restrict__ edge_off, int32_t num_pages,
                         int32_t np_cap, int32_t ne_cap, PfOut csc, PfOut csr, float* __restrict__ norm,
                         int* __restrict__ bad) {
  extern __shared__ __align__(16) int32_t pf_smem[];
  int32_t* s_src = pf_smem;                 // [ne_cap] page-local source of edge i
  int32_t* s_dst = s_src + ne_cap;          // [ne_cap]
  int32_t* slot_in = s_dst + ne_cap;        // [ne_cap] edge index per CSC position
  int32_t* slot_out = slot_in + ne_cap;     // [ne_cap]
  int32_t* ptr_in = slot_out + ne_cap;      // [np_cap + 1]
  int32_t* ptr_out = ptr_in + np_cap + 1;   // [np_cap + 1]
  int32_t* cur_in = ptr_out + np_cap + 1;   // [np_cap] counts, then fill cursors
  int32_t* cur_out = cur_in + np_cap;       // [np_cap]
  float* s_norm = reinterpret_cast<float*>(cur_out + np_cap);  // [np_cap]
  __shared__ int32_t warp_tot[2][PF_THREADS / 32];
  __shared__ int32_t carry[2];
  const int tid = threadIdx.x;
  const int page = blockIdx.x;
  const int32_t n0 = page_off[page], np = page_off[page + 1] - n0;
  const int32_t e0 = edge_off[page], ne = edge_off[page + 1] - e0;
  if (np > np_cap || ne > ne_cap) {  // cannot happen when the caller's maxima are right
    if (tid == 0) *bad = 2;
    return;
  }
  for (int i = tid; i < np; i += PF_THREADS) cur_in[i] = cur_out[i] = 0;
  __syncthreads();
  // 1. local ids + in/out degree counts
  for (int i = tid; i < ne; i += PF_THREADS) {
    const int32_t s = src[e0 + i] - n0, d = dst[e0 + i] - n0;
    const bool ok = s >= 0 && s < np && d >= 0 && d < np;
    if (!ok) *bad = 1;
    // an edge that leaves its page breaks the contract (*bad = 1, results meaningless) but must stay memory safe:
    // it is kept with its endpoints clamped to local node 0, so every count / cursor / slot below stays consistent
    const int32_t sl = ok ? s : 0, dl = ok ? d : 0;
    s_src[i] = sl;
    s_dst[i] = dl;
    atomicAdd(&cur_in[dl], 1);
    atomicAdd(&cur_out[sl], 1);
  }
  __syncthreads();
  // 2. exclusive scans of both degree arrays (block scan, PF_THREADS entries per round)
  if (tid < 2) carry[tid] = 0;
  __syncthreads();
  for (int base = 0; base < np; base += PF_THREADS) {
    const int i = base + tid;
    int32_t v[2] = {i < np ? cur_in[i] : 0, i < np ? cur_out[i] : 0};
    int32_t incl[2] = {v[0], v[1]};
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t0 = __shfl_up_sync(0xffffffffu, incl[0], o), t1 = __shfl_up_sync(0xffffffffu, incl[1], o);
      if ((tid & 31) >= o) {
        incl[0] += t0;
        incl[1] += t1;
      }
    }
    if ((tid & 31) == 31) {
      warp_tot[0][tid >> 5] = incl[0];
      warp_tot[1][tid >> 5] = incl[1];
    }
    __syncthreads();
    int32_t woff[2] = {0, 0};
    for (int k = 0; k < (tid >> 5); ++k) {
      woff[0] += warp_tot[0][k];
      woff[1] += warp_tot[1][k];
    }
    const int32_t ex0 = carry[0] + woff[0] + incl[0] - v[0], ex1 = carry[1] + woff[1] + incl[1] - v[1];
    if (i < np) {
      ptr_in[i] = ex0;
      ptr_out[i] = ex1;
    }
    __syncthreads();
    if (tid == PF_THREADS - 1) {
      carry[0] = ex0 + v[0];
      carry[1] = ex1 + v[1];
    }
    __syncthreads();
  }
  if (tid == 0) {
    ptr_in[np] = ne;
    ptr_out[np] = ne;
  }
  // 3. row pointers, normaliser; reset the cursors
  for (int i = tid; i < np; i += PF_THREADS) {
    const int32_t deg = cur_in[i];
    const float nv = deg > 0 ? 1.0f / (float)deg : 0.0f;  // 1./in_degree, inf -> 0 (models.py:75-76)
    s_norm[i] = nv;
    norm[n0 + i] = nv;
    csc.indptr[n0 + i] = e0 + ptr_in[i];
    csr.indptr[n0 + i] = e0 + ptr_out[i];
    cur_in[i] = 0;
    cur_out[i] = 0;
  }
  if (page == num_pages - 1 && tid == 0) {
    csc.indptr[n0 + np] = e0 + ne;
    csr.indptr[n0 + np] = e0 + ne;
  }
  __syncthreads();
  // 4. scatter edge indices into their rows (any order inside a row) ...
  for (int i = tid; i < ne; i += PF_THREADS) {
    const int32_t s = s_src[i], d = s_dst[i];
    slot_in[ptr_in[d] + atomicAdd(&cur_in[d], 1)] = i;
    slot_out[ptr_out[s] + atomicAdd(&cur_out[s], 1)] = i;
  }
  __syncthreads();
  // 5. ... then order every row by edge index (= stable sort of the COO): rows are short, one thread per row
  for (int r = tid; r < 2 * np; r += PF_THREADS) {
    int32_t* slot = r < np ? slot_in : slot_out;
    const int32_t* ptr = r < np ? ptr_in : ptr_out;
    const int rr = r < np ? r : r - np;
    const int b = ptr[rr], e = ptr[rr + 1];
    for (int a = b + 1; a < e; ++a) {
      const int32_t v = slot[a];
      int c = a - 1;
      while (c >= b && slot[c] > v) {
        slot[c + 1] = slot[c];
        --c;
      }
      slot[c + 1] = v;
    }
  }
  __syncthreads();
  // 6. write both formats and their packed entries
  for (int j = tid; j < ne; j += PF_THREADS) {
    const int32_t i = slot_in[j];
    const float wv = w ? __ldg(w + e0 + i) : 1.0f;
    csc.indices[e0 + j] = n0 + s_src[i];
    csc.eid[e0 + j] = e0 + i;
    csc.packed[e0 + j] = make_uint2((uint32_t)s_src[i], __float_as_uint(wv));
    const int32_t k = slot_out[j];
    float wk = w ? __ldg(w + e0 + k) : 1.0f;
    wk *= s_norm[s_dst[k]];  // source-side scale of the backward aggregation: norm[dst]
    csr.indices[e0 + j] = n0 + s_dst[k];
    csr.eid[e0 + j] = e0 + k;
    csr.packed[e0 + j] = make_uint2((uint32_t)s_dst[k], __float_as_uint(wk));
  }
}

static size_t pf_smem_bytes(int32_t np_cap, int32_t ne_cap) {
  return ((size_t)4 * ne_cap + 2 * ((size_t)np_cap + 1) + 3 * (size_t)np_cap) * 4 + 16;
}

}  // namespace gte

extern "C" size_t gte_build_page_formats_smem_bytes(int32_t max_page_nodes, int32_t max_page_edges) {
  if (max_page_nodes < 0 || max_page_edges < 0) return 0;
  const size_t b = gte::pf_smem_bytes(max_page_nodes, max_page_edges);
  return b <= 227 * 1024 ? b : 0;
}

extern "C" int gte_build_page_formats(const int32_t* src, const int32_t* dst, const float* w, const int32_t* page_off,
                                      const int32_t* edge_off, int32_t num_pages, int32_t n, int64_t e,
                                      int32_t max_page_nodes, int32_t max_page_edges, int32_t* csc_indptr,
                                      int32_t* csc_indices, int32_t* csc_eid, uint64_t* csc_packed, int32_t* csr_indptr,
                                      int32_t* csr_indices, int32_t* csr_eid, uint64_t* csr_packed, float* norm,
                                      int32_t* page_flag, int32_t* bad, gte_stream_t stream) {
  GTE_CHECK_ARG(num_pages >= 0 && n >= 0 && e >= 0 && max_page_nodes >= 0 && max_page_edges >= 0,
                "gte_build_page_formats: negative size");
  const size_t smem = gte_build_page_formats_smem_bytes(max_page_nodes, max_page_edges);
  if (smem == 0)
    return fail(GTE_ERR_UNSUPPORTED, "gte_build_page_formats: a page of %d nodes / %d edges does not fit in shared memory",
                max_page_nodes, max_page_edges);
  GTE_CHECK_ARG(page_off && edge_off && csc_indptr && csr_indptr && norm && bad && page_flag,
                "gte_build_page_formats: null argument");
  GTE_CHECK_ARG(e == 0 || (src && dst && csc_indices && csc_eid && csc_packed && csr_indices && csr_eid && csr_packed),
                "gte_build_page_formats: null edge argument");
  cudaStream_t st = as_stream(stream);
  GTE_CHECK_CUDA(cudaMemsetAsync(bad, 0, 4, st), "gte_build_page_formats(memset bad)");
  if (num_pages == 0) {
    GTE_CHECK_CUDA(cudaMemsetAsync(csc_indptr, 0, 4, st), "gte_build_page_formats(memset)");
    GTE_CHECK_CUDA(cudaMemsetAsync(csr_indptr, 0, 4, st), "gte_build_page_formats(memset)");
    return GTE_OK;
  }
  GTE_CHECK_CUDA(cudaMemsetAsync(page_flag, 0, (size_t)num_pages * 4, st), "gte_build_page_formats(memset flags)");
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(&k_build_page_formats), smem, "k_build_page_formats")) return rc;
  PfOut a{csc_indptr, csc_indices, csc_eid, reinterpret_cast<uint2*>(csc_packed)};
  PfOut b{csr_indptr, csr_indices, csr_eid, reinterpret_cast<uint2*>(csr_packed)};
  k_build_page_formats<<<num_pages, PF_THREADS, smem, st>>>(src, dst, w, page_off, edge_off, num_pages, max_page_nodes,
                                                           max_page_edges, a, b, norm, bad);
  GTE_CHECK_LAUNCH("k_build_page_formats");
  return GTE_OK;
}
