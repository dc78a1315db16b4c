// Fixed costs on the MMA-issuing warp's critical path (B200, sm_100a): what one
// tcgen05.mma issue, a tcgen05.commit, an mbarrier wait on an already-completed phase, the
// tcgen05 fences and elect.sync/__syncwarp cost the issuing warp, in cycles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_issue profiles/ubench_issue.cu
#include <cstdio>
#include <vector>
#include "../mujoco-mbrl_b200/csrc/rollout_tc.cuh"
using namespace mbrl;

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(64, 1) ubench(long long* out, int N) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 64) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (warp == 0) {
    if (lane == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc(N, true);
    const uint32_t lbo_b = (uint32_t)N * 16;
    const uint64_t bd0 = umma_desc(smem_u32(sm), lbo_b, 128);
    const uint32_t b0 = smem_u32(&bar[0]), b1 = smem_u32(&bar[1]);
    long long t[16];
    for (int rep = 0; rep < 3; ++rep) {
      const uint32_t ph = rep & 1;
      __syncwarp();
      t[0] = clock64();
      bool e = elect_one();
      __syncwarp();
      t[1] = clock64();                       // elect + syncwarp
      if (e) mma_ts(tmem, tmem + 256, bd0, idesc, 0);
      __syncwarp();
      t[2] = clock64();                       // 1 MMA issue (idle pipe)
      if (e) mma_ts(tmem, tmem + 264, bd0, idesc, 1);
      __syncwarp();
      t[3] = clock64();                       // 2nd MMA issue
      if (e) tc_commit(b0);
      __syncwarp();
      t[4] = clock64();                       // commit
      mbar_wait(b0, ph);
      t[5] = clock64();                       // wait for completion
      mbar_wait(b0, ph);
      t[6] = clock64();                       // wait on an already-completed phase (all lanes)
      tc_fence_after();
      t[7] = clock64();                       // fence::after
      tc_fence_before();
      t[8] = clock64();                       // fence::before
      bool r = mbar_test(b0, ph);
      t[9] = clock64();                       // test_wait (non-blocking)
      if (e) { mma_ts(tmem, tmem + 256, bd0, idesc, 0); mma_ts(tmem, tmem + 264, bd0, idesc, 1); mma_ts(tmem, tmem + 272, bd0, idesc, 1); mma_ts(tmem, tmem + 280, bd0, idesc, 1); tc_commit(b1); }
      __syncwarp();
      t[10] = clock64();                      // 4 MMAs + commit
      if (lane == 0) mbar_wait(b1, ph);       // single-lane wait
      __syncwarp();
      t[11] = clock64();
      if (lane == 0 && rep == 2) { for (int i = 0; i < 11; ++i) out[i] = t[i + 1] - t[i]; out[15] = r; }
    }
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 64 * 8);
  cudaFuncSetAttribute(ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  std::vector<long long> h(64);
  const char* names[] = {"elect+syncwarp", "1st MMA issue (idle pipe)", "2nd MMA issue", "commit", "wait for completion", "wait, phase already complete",
                         "fence::after_thread_sync", "fence::before_thread_sync", "test_wait", "4 MMAs + commit issue", "single-lane wait for completion"};
  for (int N : {64, 208, 240}) {
    cudaMemset(d, 0, 64 * 8);
    ubench<<<1, 64, 200 * 1024>>>(d, N);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), d, 64 * 8, cudaMemcpyDeviceToHost);
    printf("N=%d\n", N);
    for (int i = 0; i < 11; ++i) printf("  %-34s %6lld\n", names[i], h[i]);
  }
  return 0;
}
