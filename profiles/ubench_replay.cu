// Phase timeline of the replay kernel (clock64 stamps).
#define MBRL_REPLAY_PROFILE
#include <cstdio>
#include <vector>
#include "../mujoco-mbrl_b200/csrc/replay.cuh"
using namespace mbrl;
int main() {
  const int O = 17, A = 6, U = 200, D = O + A, H = 30, N = 16384;
  auto dev = [](size_t n, float v) { std::vector<float> h(n, v); float* d; cudaMalloc(&d, n * 4); cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice); return d; };
  ModelDev m{}; m.O = O; m.A = A; m.D = D; m.U = U;
  m.W1t = dev(D * U, 0.01f); m.b1 = dev(U, 0.f); m.W2t = dev(U * U, 0.01f); m.b2 = dev(U, 0.f); m.W3t = dev(U * O, 0.01f); m.b3 = dev(O, 0.f);
  m.mu_s = dev(O, 0.f); m.sd_s = dev(O, 1.f); m.mu_a = dev(A, 0.f); m.sd_a = dev(A, 1.f); m.cost_w = dev(O, 1.f); m.goal = dev(O, 0.f);
  ActionSource src{}; src.mode = MBRL_SAMPLE_GAUSSIAN; src.lo = -1; src.hi = 1;
  float* mu = dev(2 * H * A, 0.f); float* sd = dev(2 * H * A, 1.f); float* s0 = dev(O, 0.1f);
  BestEver be{1.f, 0, 5, 0}; BestEver* dbe; cudaMalloc(&dbe, sizeof(be)); cudaMemcpy(dbe, &be, sizeof(be), cudaMemcpyHostToDevice);
  float* os = dev(H * O, 0.f); float* oa = dev(H * A, 0.f);
  Shape sh{H, N, 1};
  const size_t smem = replay_smem_bytes(O, A, U, H, true);
  cudaFuncSetAttribute(replay_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int r = 0; r < 3; ++r) replay_kernel<true><<<1, kReplayThreads, smem>>>(m, src, sh, s0, mu, sd, dbe, 1, 0, 0, os, oa, nullptr);
  cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
  long long st[16]; cudaMemcpyFromSymbol(st, g_replay_stamps, sizeof(st));
  printf("smem %zu B, threads %d\nsetup (weights->smem, actions) %lld\nstep 5: L1 %lld  L2 %lld  L3 %lld  state+sync %lld\ntotal %lld cycles (%.1f per step)\n", smem, kReplayThreads,
         st[1] - st[0], st[3] - st[2], st[4] - st[3], st[5] - st[4], st[6] - st[5], st[7] - st[0], (double)(st[7] - st[1]) / H);
  {
    const RegGeom g = replay_reg_geometry(O, A, U, H);
    const size_t sm = sizeof(float) * (size_t)g.total;
    cudaFuncSetAttribute(replay_reg_kernel<5, 40>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    for (int r = 0; r < 3; ++r) replay_reg_kernel<5, 40><<<1, kRegThreads, sm>>>(m, src, sh, g, s0, mu, sd, dbe, 1, 0, os, oa, nullptr);
    e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpyFromSymbol(st, g_replay_stamps, sizeof(st));
    printf("register variant: smem %zu B, S=%d kpt2=%d\nsetup %lld\nstep 5: L1 %lld  L2 %lld (fma %lld, barrier %lld, reduce %lld)  L3 %lld\ntotal %lld cycles (%.1f per step)\n", sm, g.S, g.kpt2,
           st[1] - st[0], st[3] - st[2], st[4] - st[3], st[8] - st[3], st[9] - st[8], st[4] - st[9], st[5] - st[4], st[7] - st[0], (double)(st[7] - st[1]) / H);
  }
  return 0;
}
