// Microbenchmark: cycles per tcgen05.mma for the shapes / issue patterns of the 3xTF32 kernels (operands resident in
// shared memory, no loads in the loop).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../gnn_tableextraction_b200/csrc
// -I../../include mma_rate.cu -o mma_rate -lcuda ; run on the B200 box.
#include "gte_umma_ptx.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace gte;

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
               "l"(da), "l"(db), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
               "l"(da), "l"(db), "r"(idesc), "r"(acc)
               : "memory");
}

// mode: 0 = tf32 one accumulator, same operands; 1 = tf32 3x pattern (two A tiles, two B tiles, two accumulators);
//       2 = bf16 one accumulator; 3 = tf32 3x pattern but A tiles alternate over 3 stages (different smem)
template <bool PAIR>
__global__ void __launch_bounds__(128, 1) k_rate(int N, int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 224 * 1024 / 4; i += 128) reinterpret_cast<float*>(base)[i] = 0.f;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    if (PAIR) tmem_alloc_pair(smem_u32(&s_tmem), 512);
    else tmem_alloc(smem_u32(&s_tmem), 512);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = s_tmem;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  if (warp == 1 && lane == 0 && rank == 0) {
    const int M = PAIR ? 256 : 128;
    const uint32_t fmt = (mode == 2) ? 1u : 2u;  // bf16 : tf32
    uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (mode == 6 || mode == 7) idesc |= (1u << 15) | (1u << 16);  // both operands MN-major
    const int brows = PAIR ? N / 2 : N;
    // stage s: A_hi @ s*64K, A_lo @ +16K, B_hi @ +32K, B_lo @ +32K + brows*128
    auto desc = [&](int s, int which) {
      uint32_t off = s * 64 * 1024;
      if (which == 1) off += 16 * 1024;
      if (which == 2) off += 32 * 1024;
      if (which == 3) off += 32 * 1024 + brows * 128;
      if (mode == 6 || mode == 7) return make_desc_mn_sw128_32b(smem_u32(base + off), 2048);  // 16-row boxes
      return make_desc_k_sw128(smem_u32(base + off));
    };
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int s = (mode == 3) ? it % 3 : 0;
      const uint64_t dah = desc(s, 0), dal = desc(s, 1), dbh = desc(s, 2), dbl = desc(s, 3);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t adv = (uint64_t)((k * 32) >> 4);
        if (mode == 0) {
          for (int j = 0; j < 3; ++j) { if (PAIR) umma_tf32_pair(tm, dah + adv, dbh + adv, idesc, 1u); else umma_tf32(tm, dah + adv, dbh + adv, idesc, 1u); }
        } else if (mode == 2) {
          for (int j = 0; j < 3; ++j) { if (PAIR) umma_f16_pair(tm, dah + adv, dbh + adv, idesc, 1u); else umma_f16(tm, dah + adv, dbh + adv, idesc, 1u); }
        } else if (mode == 4) {  // 3x pattern, grouped by accumulator: the four main MMAs of the k-block, then the eight cross MMAs
          if (k == 0) {
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t a2 = (uint64_t)((kk * 32) >> 4);
              if (PAIR) umma_tf32_pair(tm, dah + a2, dbh + a2, idesc, 1u); else umma_tf32(tm, dah + a2, dbh + a2, idesc, 1u);
            }
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t a2 = (uint64_t)((kk * 32) >> 4);
              if (PAIR) { umma_tf32_pair(tm + 256, dal + a2, dbh + a2, idesc, 1u); umma_tf32_pair(tm + 256, dah + a2, dbl + a2, idesc, 1u); }
              else { umma_tf32(tm + 256, dal + a2, dbh + a2, idesc, 1u); umma_tf32(tm + 256, dah + a2, dbl + a2, idesc, 1u); }
            }
          }
        } else if (mode == 6) {  // MN-major operands (weight-gradient kernel), 3x pattern; K step = 1024 B
          const uint64_t a6 = (uint64_t)(((k & 1) * 1024) >> 4);
          if (PAIR) { umma_tf32_pair(tm + 256, dal + a6, dbh + a6, idesc, 1u); umma_tf32_pair(tm + 256, dah + a6, dbl + a6, idesc, 1u); umma_tf32_pair(tm, dah + a6, dbh + a6, idesc, 1u); }
          else { umma_tf32(tm + 256, dal + a6, dbh + a6, idesc, 1u); umma_tf32(tm + 256, dah + a6, dbl + a6, idesc, 1u); umma_tf32(tm, dah + a6, dbh + a6, idesc, 1u); }
        } else if (mode == 7) {  // MN-major, one accumulator, same operands
          for (int j = 0; j < 3; ++j) { if (PAIR) umma_tf32_pair(tm, dah, dbh, idesc, 1u); else umma_tf32(tm, dah, dbh, idesc, 1u); }
        } else if (mode == 5) {  // same accumulator for everything, operands alternate as in the 3x pattern
          if (PAIR) {
            umma_tf32_pair(tm, dal + adv, dbh + adv, idesc, 1u); umma_tf32_pair(tm, dah + adv, dbl + adv, idesc, 1u); umma_tf32_pair(tm, dah + adv, dbh + adv, idesc, 1u);
          } else {
            umma_tf32(tm, dal + adv, dbh + adv, idesc, 1u); umma_tf32(tm, dah + adv, dbl + adv, idesc, 1u); umma_tf32(tm, dah + adv, dbh + adv, idesc, 1u);
          }
        } else {
          if (PAIR) {
            umma_tf32_pair(tm + 256, dal + adv, dbh + adv, idesc, 1u);
            umma_tf32_pair(tm + 256, dah + adv, dbl + adv, idesc, 1u);
            umma_tf32_pair(tm, dah + adv, dbh + adv, idesc, 1u);
          } else {
            umma_tf32(tm + 256, dal + adv, dbh + adv, idesc, 1u);
            umma_tf32(tm + 256, dah + adv, dbl + adv, idesc, 1u);
            umma_tf32(tm, dah + adv, dbh + adv, idesc, 1u);
          }
        }
      }
    }
    if (PAIR) umma_commit_pair(smem_u32(&bar)); else umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tm, 512); else tmem_dealloc(tm, 512);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  const size_t smem = 226 * 1024;
  cudaFuncSetAttribute(k_rate<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_rate<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 200;
  const char* names[] = {"tf32 1acc same-operands", "tf32 3x pattern", "bf16 1acc", "tf32 3x pattern, 3 stages", "tf32 3x grouped by acc", "tf32 3x operands, 1 acc", "tf32 MN-major 3x pattern", "tf32 MN-major 1acc"};
  for (int pair = 0; pair < 2; ++pair)
    for (int N : {224, 256, 64})
      for (int mode = 0; mode < 8; ++mode) {
        if (pair && (N / 2) % 8) continue;
        cudaMemset(d, 0, 148 * 8);
        if (pair) {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
          cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
          cfg.attrs = &at; cfg.numAttrs = 1;
          cudaLaunchKernelEx(&cfg, k_rate<true>, N, mode, iters, d);
        } else {
          k_rate<false><<<148, 128, smem>>>(N, mode, iters, d);
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(148);
        cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
        long long mx = 0, mn = 1LL << 60; int cnt = 0;
        for (auto v : h) if (v > 0) { mx = v > mx ? v : mx; mn = v < mn ? v : mn; ++cnt; }
        printf("%s N=%3d %-28s cycles/MMA min %.1f max %.1f (%d issuers)  => per 12-MMA k-block %.0f\n", pair ? "pair  " : "single", N, names[mode],
               (double)mn / (iters * 12), (double)mx / (iters * 12), cnt, (double)mx / iters);
      }
  return 0;
}
