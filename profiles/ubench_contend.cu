// Does tcgen05.mma throughput drop while many warps do TMEM load/convert/store round trips
// (the hidden epilogue's traffic)?  W traffic warps (W/4 per TMEM lane quarter) + 1 MMA warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_contend profiles/ubench_contend.cu
#include <cstdio>
#include <vector>
#include "../mujoco-mbrl_b200/csrc/rollout_tc.cuh"
using namespace mbrl;

__global__ void __launch_bounds__(1024, 1) ubench(long long* out, int n_mma, int N, int W, int trips, int mode) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ int go;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mma_warp = W;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (warp == mma_warp) {
    if (lane == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == mma_warp) {
    const uint32_t idesc = umma_idesc(N, true);
    const uint32_t lbo_b = (uint32_t)N * 16;
    const uint64_t bd0 = umma_desc(smem_u32(sm), lbo_b, 128);
    const uint64_t step_b = (2 * lbo_b) >> 4;
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      if (elect_one()) {
        uint64_t bd = bd0;
        for (int i = 0; i < n_mma; ++i) {
          mma_ts(tmem, tmem + 256 + 8 * (i & 15), bd, idesc, i > 0);
          bd += step_b; if ((i & 7) == 7) bd = bd0;
        }
        tc_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), rep & 1);
      long long t2 = clock64();
      if (lane == 0) out[rep] = t2 - t0;
    }
  } else if (warp < W) {
    // traffic: the epilogue's unit (ld 16 columns, relu-pack, st 8 columns) on columns the MMA does not touch
    uint32_t v[32], pk[16];
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 400 + 16 * ((warp >> 2) & 3);
    long long t0 = clock64();
    for (int it = 0; it < trips; ++it) {
      if (mode != 2) { tmem_ld16(base, v); tmem_ld_wait(); }
#pragma unroll
      for (int i = 0; i < 8; ++i) pk[i] = pack_relu<true>(v[2 * i], v[2 * i + 1]);
      if (mode != 1) { tmem_st8(base, pk); tmem_st_wait(); }
      if (mode == 2) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = pk[i & 7] + it;
      }
    }
    long long t1 = clock64();
    if (lane == 0) out[8 + warp] = (t1 - t0) / (trips > 0 ? trips : 1);
    if (pk[3] == 0x12345u) out[63] = v[5];
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == mma_warp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 64 * 8);
  cudaFuncSetAttribute(ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  std::vector<long long> h(64);
  const char* modes[] = {"ld+cvt+st", "ld+cvt only", "cvt+st only"};
  for (int N : {208})
    for (int mode = 0; mode < 3; ++mode)
      for (int W : {0, 4, 8, 16}) {
        if (W == 0 && mode > 0) continue;
        cudaMemset(d, 0, 64 * 8);
        ubench<<<1, (W + 1) * 32, 200 * 1024>>>(d, 52, N, W, W ? 60 : 0, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d, 64 * 8, cudaMemcpyDeviceToHost);
        printf("N=%d 52 TS MMAs, %2d traffic warps (%s): %5lld cycles (%.1f per MMA); unit round trip: warp0 %lld, last %lld cycles\n", N, W, modes[mode], h[2],
               (double)h[2] / 52, W ? h[8] : 0, W ? h[8 + W - 1] : 0);
      }
  return 0;
}
