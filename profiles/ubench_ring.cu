// Issue-loop structure of a weight-streaming tcgen05 consumer (B200, sm_100a): a producer warp
// TMA-bulk-copies B tiles (N=256, K=16: 8 KB) from an L2-resident image through a shared-memory ring,
// the MMA warp issues one TS MMA per tile.  Which loop structure keeps the tensor pipe busy?
//   V0  warp-converged waits, elect + commit + __syncwarp per stage (round-2 first version)
//   V1  one elected lane runs the whole loop
//   V2  V1 + the next stage's full-barrier wait is taken BEFORE the last MMA of the current stage
//   V3  V1 + the wait for stage s+1 right after the FIRST MMA of stage s
// for 2 and 4 tiles per stage, on 1 CTA and on 148 CTAs (all SMs streaming at once).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_ring profiles/ubench_ring.cu
#include <cstdio>
#include <vector>
#include "../mujoco-mbrl_b200/csrc/rollout_tc.cuh"
using namespace mbrl;

constexpr int kTile = 8192;

template <int V, int TPS, int F>
__global__ void __launch_bounds__((F & 1) ? 832 : 96, 1) ubench(const uint8_t* __restrict__ img, int tiles_total, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  constexpr int SB = TPS * kTile;
  constexpr int S = 96 * 1024 / SB;  // 96 KB ring
  __shared__ uint64_t bar[2 * 16 + 1];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(&bar[0]), bar_empty = smem_u32(&bar[16]), bar_done = smem_u32(&bar[32]);
  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < 16; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
      mbar_init(bar_done, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t ring = smem_u32(sm) + ((F & 2) ? 105344u : 0u);  // F&2: ring at the real kernel's offset in a 227 KB allocation
  const int stages_total = tiles_total / TPS;
  if (warp == 1) {
    if (elect_one()) {
      uint32_t st = 0, ph = 0;
      for (int s = 0; s < stages_total; ++s) {
        mbar_wait(bar_empty + 8 * st, ph ^ 1);
        mbar_arrive_expect_tx(bar_full + 8 * st, SB);
        bulk_g2s(ring + st * SB, img + (size_t)(s % 40) * SB, SB, bar_full + 8 * st);
        if (++st == S) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 0) {
    const uint32_t idesc = umma_idesc(256, true);
    const uint64_t bd0 = umma_desc(ring, 256 * 16, 128);
    uint32_t st = 0, ph = 0;
    const long long t0 = clock64();
    if (V == 0) {
      for (int s = 0; s < stages_total; ++s) {
        mbar_wait(bar_full + 8 * st, ph);
        tc_fence_after();
        if (elect_one()) {
          for (int i = 0; i < TPS; ++i) mma_ts(tmem + 256, tmem + 8 * ((s * TPS + i) & 15), bd0 + ((st * SB + i * kTile) >> 4), idesc, 1);
          tc_commit(bar_empty + 8 * st);
          if (s + 1 == stages_total) tc_commit(bar_done);
        }
        __syncwarp();
        if (++st == S) { st = 0; ph ^= 1; }
      }
    } else {
      if (elect_one()) {
        mbar_wait(bar_full, 0);
        tc_fence_after();
        for (int s = 0; s < stages_total; ++s) {
          uint32_t nst = st + 1, nph = ph;
          if (nst == S) { nst = 0; nph ^= 1; }
#pragma unroll
          for (int i = 0; i < TPS; ++i) {
            if (V == 2 && i == TPS - 1 && s + 1 < stages_total) { mbar_wait(bar_full + 8 * nst, nph); tc_fence_after(); }
            mma_ts(tmem + 256, tmem + 8 * ((s * TPS + i) & 15), bd0 + ((st * SB + i * kTile) >> 4), idesc, 1);
            if (V == 3 && i == 0 && s + 1 < stages_total) { mbar_wait(bar_full + 8 * nst, nph); tc_fence_after(); }
          }
          tc_commit(bar_empty + 8 * st);
          if (s + 1 == stages_total) tc_commit(bar_done);
          if (V == 1 && s + 1 < stages_total) { mbar_wait(bar_full + 8 * nst, nph); tc_fence_after(); }
          st = nst; ph = nph;
        }
      }
      __syncwarp();
    }
    mbar_wait(bar_done, 0);
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  }
  if (warp >= 2) mbar_wait(bar_done, 0);  // F&1: 24 more warps parked on an mbarrier for the whole run, like the idle roles of the real kernel
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int V, int TPS, int F = 0>
void run(const uint8_t* img, long long* d, int grid) {
  const int tiles = 3200;
  const int smem = (F & 2) ? 232448 - 1024 : 96 * 1024 + 1024;
  cudaFuncSetAttribute(ubench<V, TPS, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<long long> h(grid);
  for (int rep = 0; rep < 2; ++rep) {
    ubench<V, TPS, F><<<grid, (F & 1) ? 832 : 96, smem>>>(img, tiles, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h.data(), d, grid * 8, cudaMemcpyDeviceToHost);
  long long mx = 0, sum = 0;
  for (long long v : h) { mx = v > mx ? v : mx; sum += v; }
  printf("V%d F%d tiles/stage %d grid %3d : %.1f cycles per MMA (mean over CTAs), %.1f (slowest CTA)\n", V, F, TPS, grid,
         (double)sum / grid / tiles, (double)mx / tiles);
}

int main() {
  uint8_t* img; cudaMalloc(&img, 40 * 4 * kTile); cudaMemset(img, 0, 40 * 4 * kTile);
  long long* d; cudaMalloc(&d, 256 * 8);
  for (int grid : {1, 148}) {
    run<0, 2>(img, d, grid); run<1, 2>(img, d, grid); run<2, 2>(img, d, grid); run<3, 2>(img, d, grid);
    run<0, 4>(img, d, grid); run<1, 4>(img, d, grid); run<2, 4>(img, d, grid); run<3, 4>(img, d, grid);
    run<0, 2, 1>(img, d, grid); run<0, 2, 2>(img, d, grid); run<0, 2, 3>(img, d, grid); run<1, 2, 3>(img, d, grid);
  }
  return 0;
}
