// Microbenchmark: how many bytes per clock can ONE SM pull through TMA 2-D box loads of 128-byte rows (SWIZZLE_128B,
// the operand-tile shape of the 3xTF32 kernels), as a function of box height, requests per stage and ring depth?
// One producer thread issues, one consumer thread frees the stage as soon as the bytes landed: nothing else runs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../gnn_tableextraction_b200/csrc -I../../include
//        tma_rate.cu -o tma_rate -lcuda ; run on the B200 box:  ./tma_rate
#include "gte_umma_ptx.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
namespace gte {
char* err_buf() { static char b[256]; return b; }
int fail(int code, const char* fmt, ...) { printf("fail: %s\n", fmt); return code; }
}
using namespace gte;

struct Args {
  CUtensorMap tm;
  int box_rows, reqs, stages, iters, nrow_boxes, ncol_boxes;
};

__global__ void __launch_bounds__(64, 1) k_tma(const __grid_constant__ Args P, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[8], empty[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int box_bytes = P.box_rows * 128;
  const int stage_bytes = P.reqs * box_bytes;
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    fence_barrier_init();
  }
  __syncthreads();
  // box walk without divisions in the loop: this CTA's column box is fixed, its row box advances by a stride
  const int cb = blockIdx.x % P.ncol_boxes;
  int rb = (int)(((long long)blockIdx.x * 7919) % P.nrow_boxes);
  const int rstep = 37 % P.nrow_boxes;
  if (warp == 0 && lane == 0) {
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < P.iters; ++it) {
      mbar_wait(smem_u32(&empty[s]), ph ^ 1);
      mbar_expect_tx(smem_u32(&full[s]), (uint32_t)stage_bytes);
      for (int r = 0; r < P.reqs; ++r) {
        tma_load_2d(smem_u32(base + s * stage_bytes + r * box_bytes), &P.tm, smem_u32(&full[s]), cb * 32, rb * P.box_rows);
        rb += rstep; if (rb >= P.nrow_boxes) rb -= P.nrow_boxes;
      }
      if (++s == P.stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    int s = 0; uint32_t ph = 0;
    long long t0 = 0;
    for (int it = 0; it < P.iters; ++it) {
      mbar_wait(smem_u32(&full[s]), ph);
      if (it == P.stages) t0 = clock64();
      mbar_arrive(smem_u32(&empty[s]));
      if (++s == P.stages) { s = 0; ph ^= 1; }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

int main(int argc, char** argv) {
  const int ld = 224;
  struct Cfg { long long rows; int box_rows, reqs, stages; };
  std::vector<Cfg> cfgs;
  for (long long rows : {16384LL, 1228800LL})
    for (int br : {16, 64, 128, 256})
      for (int reqs : {1, 2, 4})
        for (int st : {2, 4, 8}) {
          if ((long long)br * 128 * reqs * st > 200 * 1024) continue;
          cfgs.push_back({rows, br, reqs, st});
        }
  float* d = nullptr;
  cudaMalloc(&d, 1228800LL * ld * 4);
  {  // non-trivial contents (all-zero lines could be compressed on the way)
    std::vector<float> h(1 << 20);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)((i * 2654435761u) >> 8) * 1e-3f;
    for (long long off = 0; off < 1228800LL * ld; off += (long long)h.size()) {
      const long long n = std::min<long long>((long long)h.size(), 1228800LL * ld - off);
      cudaMemcpy(d + off, h.data(), n * 4, cudaMemcpyHostToDevice);
    }
  }
  long long* out; cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  printf("rows box_rows reqs stages | KB/stage in-flight KB | B/clk/SM (median) | GB/s aggregate\n");
  for (auto& c : cfgs) {
    Args a{};
    if (make_tmap_2d(&a.tm, d, c.rows, ld, ld, 32, c.box_rows)) return 1;
    a.box_rows = c.box_rows; a.reqs = c.reqs; a.stages = c.stages;
    a.nrow_boxes = (int)(c.rows / c.box_rows); a.ncol_boxes = ld / 32;
    const long long stage_bytes = (long long)c.box_rows * 128 * c.reqs;
    a.iters = (int)((48LL << 20) / stage_bytes);  // 48 MB per SM
    if (a.iters < 64) a.iters = 64;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_tma<<<148, 64, 205 * 1024>>>(a, out);  // warm-up (L2 for the small matrix)
    cudaEventRecord(e0);
    k_tma<<<148, 64, 205 * 1024>>>(a, out);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(err)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), out, 148 * 8, cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    const double bytes_per_sm = (double)stage_bytes * (a.iters - c.stages);
    printf("%8lld %4d %2d %2d | %6.1f %6.1f | %6.2f | %8.1f\n", c.rows, c.box_rows, c.reqs, c.stages, stage_bytes / 1024.0,
           stage_bytes * c.stages / 1024.0, bytes_per_sm / (double)h[74], 148.0 * stage_bytes * a.iters / (ms * 1e6));
  }
  return 0;
}
