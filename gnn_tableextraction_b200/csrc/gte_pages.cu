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
