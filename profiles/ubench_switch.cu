// What does it cost to change the tcgen05.mma instruction shape / accumulator / operand source
// between consecutive MMAs of one stream?  (B200, sm_100a)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_switch profiles/ubench_switch.cu
#include <cstdio>
#include <vector>
#include "../mujoco-mbrl_b200/csrc/rollout_tc.cuh"
using namespace mbrl;

struct Op { int ts, N, dcol, count; };  // count MMAs of this kind in a row

__global__ void __launch_bounds__(64, 1) ubench(long long* out, const Op* ops, int nops) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 64) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (warp == 0) {
    if (lane == 0) { mbar_init(smem_u32(&bar), 1); for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bar2[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    const uint32_t lbo_b = 256 * 16, lbo_a = 128 * 16;
    const uint64_t bd0 = umma_desc(smem_u32(sm), lbo_b, 128), ad0 = umma_desc(smem_u32(sm + 140 * 1024), lbo_a, 128);
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      if (elect_one()) {
        int first = 1;
        for (int o = 0; o < nops; ++o) {
          const Op op = ops[o];
          if (op.N == 0) { tc_commit(smem_u32(&bar2[o & 7])); continue; }  // mid-stream commit on another barrier
          const uint32_t idesc = umma_idesc(op.N, true);
          for (int i = 0; i < op.count; ++i) {
            if (op.ts) mma_ts(tmem + op.dcol, tmem + 448 + 8 * (i & 7), bd0 + (uint64_t)(i & 7) * 32, idesc, first ? 0u : 1u);
            else mma_ss(tmem + op.dcol, ad0, bd0 + (uint64_t)(i & 7) * 32, idesc, first ? 0u : 1u);
            first = 0;
          }
        }
        tc_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), rep & 1);
      long long t2 = clock64();
      if (lane == 0) out[rep] = t2 - t0;
    }
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 64 * 8);
  Op* dops; cudaMalloc(&dops, 64 * sizeof(Op));
  cudaFuncSetAttribute(ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Case { const char* name; std::vector<Op> ops; };
  std::vector<Case> cases = {
    {"13 x TS N=208 (one accumulator)", {{1, 208, 0, 13}}},
    {"26 x TS N=208 (one accumulator)", {{1, 208, 0, 26}}},
    {"13 TS N=208 @D0, then 13 TS N=208 @D208 (accumulator switch)", {{1, 208, 0, 13}, {1, 208, 208, 13}}},
    {"8 TS N=208, 5 TS N=64 @D0, 5 TS N=144 @D64 (tail split)", {{1, 208, 0, 8}, {1, 64, 0, 5}, {1, 144, 64, 5}}},
    {"13 TS N=64 @D0, 13 TS N=144 @D64 (full N split)", {{1, 64, 0, 13}, {1, 144, 64, 13}}},
    {"13 TS N=208 @D0, 13 TS N=32 @D208 (separate y pass)", {{1, 208, 0, 13}, {1, 32, 208, 13}}},
    {"1 SS N=240, then 13 TS N=240 (action K-step first)", {{0, 240, 0, 1}, {1, 240, 0, 13}}},
    {"14 x TS N=240", {{1, 240, 0, 14}}},
    {"13 TS N=208, COMMIT, 13 TS N=208 @D208", {{1, 208, 0, 13}, {1, 0, 0, 0}, {1, 208, 208, 13}}},
    {"13 TS N=208, COMMIT, 13 TS N=208 same accumulator", {{1, 208, 0, 13}, {1, 0, 0, 0}, {1, 208, 0, 13}}},
    {"8 TS N=208, 5 TS N=64 @D0, COMMIT, 5 TS N=144 @D64 (tail split as in the kernel)", {{1, 208, 0, 8}, {1, 64, 0, 5}, {1, 0, 0, 0}, {1, 144, 64, 5}}},
    {"1 SS N=240, COMMIT, 13 TS N=240 (production GEMM-A)", {{0, 240, 0, 1}, {1, 0, 0, 0}, {1, 240, 0, 13}}},
    {"alternating TS N=64@D0 / N=144@D64, 13 pairs", {}},
  };
  for (int i = 0; i < 13; ++i) { cases.back().ops.push_back({1, 64, 0, 1}); cases.back().ops.push_back({1, 144, 64, 1}); }
  std::vector<long long> h(8);
  for (auto& c : cases) {
    cudaMemcpy(dops, c.ops.data(), c.ops.size() * sizeof(Op), cudaMemcpyHostToDevice);
    ubench<<<1, 64, 200 * 1024>>>(d, dops, (int)c.ops.size());
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), d, 64, cudaMemcpyDeviceToHost);
    printf("%-70s : %6lld cycles\n", c.name, h[2]);
  }
  return 0;
}
