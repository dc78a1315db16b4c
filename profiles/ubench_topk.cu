// Phase timeline of the top-k kernel (clock64 stamps; compile with -DMBRL_TOPK_PROFILE).
#define MBRL_TOPK_PROFILE
#include <cstdio>
#include <vector>
#include <random>
#include "../mujoco-mbrl_b200/csrc/select.cuh"
using namespace mbrl;
int main() {
  const int n = 16384, k = 1638;
  std::vector<float> h(n); std::mt19937 g(1); std::normal_distribution<float> d(300.f, 40.f);
  for (auto& x : h) x = d(g);
  float* dc; int* di; cudaMalloc(&dc, n * 4); cudaMalloc(&di, k * 4);
  cudaMemcpy(dc, h.data(), n * 4, cudaMemcpyHostToDevice);
  const size_t smem = 4 * (size_t)select_padded(n);
  cudaFuncSetAttribute(topk_select_kernel<true, kSelPlain>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 3; ++rep) topk_select_kernel<true, kSelPlain><<<1, kSelectThreads, smem>>>(dc, n, k, di, nullptr, nullptr, nullptr, 0, SelShard{});
  cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
  long long st[32]; cudaMemcpyFromSymbol(st, g_topk_stamps, sizeof(st));
  const char* names[32] = {};
  printf("stage+minmax %lld\n", st[1] - st[0]);
  for (int r = 0; r < 5; ++r) if (st[4 + 3 * r] > st[0]) printf("round %d: hist %lld  scan %lld\n", r, st[3 + 3 * r] - st[2 + 3 * r], st[4 + 3 * r] - st[3 + 3 * r]);
  printf("compaction: load+count %lld  scan %lld  smem scatter %lld  barrier %lld  write-out %lld\n", st[23] - st[20], st[24] - st[23], st[25] - st[24], st[26] - st[25], st[21] - st[26]);
  printf("compaction %lld  argmin %lld  total %lld cycles\n", st[21] - st[20], st[22] - st[21], st[22] - st[0]);
  return 0;
}
