// Micro-benchmarks of the tcgen05 building blocks used by the rollout kernels (B200, sm_100a):
// back-to-back MMA cost by shape / operand source, TMEM load/store round trips, relu+pack.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench profiles/ubench_tcgen05.cu
#include <cstdio>
#include <vector>
#include "../mujoco-mbrl_b200/csrc/rollout_tc.cuh"
using namespace mbrl;

// one CTA, 160 threads: warps 0-3 TMEM load/store tests, warp 4 MMA issue
__global__ void __launch_bounds__(160, 1) ubench(long long* out, int n_mma, int N, int ts_mode, int swz_unused) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 160) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;  // 1.0h
  if (warp == 4) {
    if (lane == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 4) {
    const uint32_t idesc = umma_idesc(N, true);
    const uint32_t lbo_b = (uint32_t)N * 16, lbo_a = 128 * 16;
    const uint64_t bd0 = umma_desc(smem_u32(sm), lbo_b, 128), ad0 = umma_desc(smem_u32(sm + 120 * 1024), lbo_a, 128);
    const uint64_t step_b = (2 * lbo_b) >> 4;
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      if (elect_one()) {
        uint64_t bd = bd0;
        for (int i = 0; i < n_mma; ++i) {
          if (ts_mode) mma_ts(tmem, tmem + 256 + 8 * (i & 15), bd, idesc, i > 0);
          else mma_ss(tmem, ad0, bd, idesc, i > 0);
          bd += step_b; if ((i & 7) == 7) bd = bd0;
        }
        tc_commit(smem_u32(&bar));
      }
      __syncwarp();
      long long t1 = clock64();
      mbar_wait(smem_u32(&bar), rep & 1);
      long long t2 = clock64();
      if (lane == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
    }
  } else if (swz_unused) {
    // contention mode: hammer TMEM with load/convert/store round trips while warp 4 times its MMAs
    uint32_t v[32], pk[16];
    const uint32_t base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int it = 0; it < swz_unused; ++it) {
      tmem_ld32(base + 300 + 32 * (it & 3), v); tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = pack_relu<true>(v[2 * i], v[2 * i + 1]);
      tmem_st16(base + 300 + 32 * (it & 3), pk); tmem_st_wait();
    }
    if (pk[3] == 0x12345u) out[63] = v[5];
  } else {
    // TMEM round trips by one warp (others idle) and by four warps
    uint32_t v[32], pk[16];
    const uint32_t base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int rep = 0; rep < 2; ++rep) {
      __syncwarp();
      long long t0 = clock64();
      tmem_ld32(base, v); tmem_ld_wait();
      long long t1 = clock64();
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = pack_relu<true>(v[2 * i], v[2 * i + 1]);
      long long t2 = clock64();
      tmem_st16(base + 300, pk); tmem_st_wait();
      long long t3 = clock64();
      tc_fence_before();
      long long t4 = clock64();
      if (lane == 0 && rep == 1) { long long* o = out + 8 + warp * 4; o[0] = t1 - t0; o[1] = t2 - t1; o[2] = t3 - t2; o[3] = t4 - t3; }
      if (pk[3] == 0x12345u) out[63] = v[5];
    }
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 64 * 8);
  cudaFuncSetAttribute(ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  std::vector<long long> h(64);
  printf("mode  N  n_mma : issue_cycles  done_cycles  (per MMA)\n");
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {32, 64, 128, 192, 208, 224, 240, 256})
      for (int n : {1, 2, 4, 13, 26}) {
        cudaMemset(d, 0, 64 * 8);
        ubench<<<1, 160, 200 * 1024>>>(d, n, N, ts, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d, 64 * 8, cudaMemcpyDeviceToHost);
        printf("%s %4d %3d : %6lld %6lld  (%.1f)\n", ts ? "TS" : "SS", N, n, h[4], h[5], (double)h[5] / n);
      }
  printf("-- MMA cost with 4 warps doing TMEM ld32/pack/st16 round trips concurrently --\n");
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {208, 240})
      for (int traffic : {0, 40, 400}) {
        cudaMemset(d, 0, 64 * 8);
        ubench<<<1, 160, 200 * 1024>>>(d, 26, N, ts, traffic);
        cudaDeviceSynchronize();
        cudaMemcpy(h.data(), d, 64 * 8, cudaMemcpyDeviceToHost);
        printf("%s N=%d 26 MMAs, %3d concurrent round trips/warp: done %lld (%.1f per MMA)\n", ts ? "TS" : "SS", N, traffic, h[5], (double)h[5] / 26);
      }
  ubench<<<1, 160, 200 * 1024>>>(d, 1, 64, 1, 0); cudaDeviceSynchronize(); cudaMemcpy(h.data(), d, 64 * 8, cudaMemcpyDeviceToHost);
  printf("TMEM per warp (ld32+wait, 16x relu-pack, st16+wait, fence): ");
  for (int w = 0; w < 4; ++w) printf("[%lld %lld %lld %lld] ", h[8 + w * 4], h[9 + w * 4], h[10 + w * 4], h[11 + w * 4]);
  printf("\n");
  return 0;
}
