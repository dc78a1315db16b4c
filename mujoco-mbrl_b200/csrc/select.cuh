// Elite selection (segmented radix-select top-k) and mean/std refit kernels.
//
// The reference has no CEM; its only selection is
// np.argmin(trajectory_costs) (src/mbrl/planners.py:184: first minimum on ties).  The
// top-k here is defined to be consistent with it: the k smallest costs, ties toward the
// lower index == np.argsort(costs, kind="stable")[:k] as a set, emitted in ascending
// index order.  Integer/index work: results are bit-exact against the oracle.
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace mbrl {

// Monotone float -> uint32 key: ascending key order == ascending float order, -0 == +0,
// every NaN sorts last (numpy's convention).
__device__ __forceinline__ uint32_t cost_key(float c) {
  if (c != c) return 0xFFFFFFFFu;
  const uint32_t u = __float_as_uint(c + 0.0f);  // -0 -> +0
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int kSelectThreads = 1024;

struct BestEver {  // per environment, device resident
  float cost;
  int iteration;
  int index;
  int pad;
};

// One CTA per segment (environment).  The float costs are mapped to monotone uint32 keys and
// staged once in shared memory (coalesced, loads in flight together) while the block min/max is
// reduced.  The k-th smallest key T is then found by an ADAPTIVE radix select: every round
// histograms the keys that are still in range into 1024 equal-width buckets of the CURRENT key
// range [lo, hi] (so the candidates spread over the bins instead of piling onto the few leading
// bit patterns that the costs of one population share), a block scan locates the bucket holding
// the k-th key, and the range shrinks 1024x; it ends when the bucket width is 1 (<= 4 rounds).
// A final index-ordered compaction (ballot + warp-shuffle scans) emits every key < T plus the
// first `take_eq` keys == T.
//   best (nullable):      (min cost, ., argmin) of this launch per segment
//   best_ever (nullable): updated when this launch's minimum is strictly smaller
//                         (earlier iteration wins ties)
// STAGED=false re-reads the costs from global memory (segments too long for shared memory).
constexpr int kSelectStageMax = 49152;  // keys staged in shared memory: 192 KB
constexpr int kSelectBins = 1024;

template <bool STAGED>
__global__ void __launch_bounds__(kSelectThreads)
topk_select_kernel(const float* __restrict__ costs, int n, int k, int* __restrict__ elite_idx,
                   float* __restrict__ elite_cost, MbrlPlanInfo* __restrict__ best,
                   BestEver* __restrict__ best_ever, int iteration) {
  extern __shared__ __align__(16) uint32_t sel_smem[];
  uint32_t* keys = sel_smem;  // [n] when STAGED
  __shared__ uint32_t hist[kSelectBins];
  __shared__ uint32_t s_lo, s_hi, s_remaining;
  __shared__ uint32_t warp_less[32], warp_eq[32];
  __shared__ unsigned long long warp_min[32];
  __shared__ uint32_t s_base_less, s_base_eq;

  const int seg = blockIdx.x;
  const float* c = costs + (long long)seg * n;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  auto key_at = [&](int i) -> uint32_t { return STAGED ? keys[i] : cost_key(__ldg(c + i)); };

  // ---- stage + min/max ----
  uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
#pragma unroll 8
  for (int i = t; i < n; i += kSelectThreads) {
    const uint32_t key = cost_key(__ldg(c + i));
    if (STAGED) keys[i] = key;
    kmin = min(kmin, key);
    kmax = max(kmax, key);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xFFFFFFFFu, kmin, d));
    kmax = max(kmax, __shfl_xor_sync(0xFFFFFFFFu, kmax, d));
  }
  if (lane == 0) { warp_less[warp] = kmin; warp_eq[warp] = kmax; }
  __syncthreads();
  if (warp == 0) {
    kmin = warp_less[lane]; kmax = warp_eq[lane];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      kmin = min(kmin, __shfl_xor_sync(0xFFFFFFFFu, kmin, d));
      kmax = max(kmax, __shfl_xor_sync(0xFFFFFFFFu, kmax, d));
    }
    if (lane == 0) { s_lo = kmin; s_hi = kmax; s_remaining = (uint32_t)k; s_base_less = 0; s_base_eq = 0; }
  }
  __syncthreads();

  // ---- adaptive radix select ----
  for (int round = 0; round < 5; ++round) {
    const uint32_t lo = s_lo, hi = s_hi, rem = s_remaining;
    const uint32_t width = (uint32_t)(((unsigned long long)(hi - lo)) / kSelectBins) + 1u;  // bucket width
    hist[t] = 0;
    __syncthreads();
#pragma unroll 4
    for (int i = t; i < n; i += kSelectThreads) {
      const uint32_t key = key_at(i);
      if (key >= lo && key <= hi) atomicAdd(&hist[(key - lo) / width], 1u);
    }
    __syncthreads();
    // block-wide inclusive scan of the 1024 bins (one bin per thread)
    const uint32_t mine = hist[t];
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) warp_less[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const uint32_t tot = warp_less[lane];
      uint32_t wi = tot;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, wi, d);
        if (lane >= d) wi += v;
      }
      warp_eq[lane] = wi - tot;  // exclusive prefix of the warp totals
    }
    __syncthreads();
    incl += warp_eq[warp];
    const uint32_t excl = incl - mine;
    if (rem > excl && rem <= incl) {  // exactly one bin holds the rem-th smallest in-range key
      const uint32_t nlo = lo + (uint32_t)t * width;
      const unsigned long long nhi = (unsigned long long)nlo + width - 1ull;
      s_lo = nlo;
      s_hi = nhi < (unsigned long long)hi ? (uint32_t)nhi : hi;
      s_remaining = rem - excl;
    }
    __syncthreads();
    if (width == 1u) break;  // uniform: the bucket is a single key value
  }
  const uint32_t T = s_lo;               // k-th smallest key
  const uint32_t take_eq = s_remaining;  // how many keys == T belong to the elite set

  // ---- index-ordered compaction + argmin ----
  unsigned long long my_min = ~0ull;
  const int rounds = (n + kSelectThreads - 1) / kSelectThreads;
  for (int r = 0; r < rounds; ++r) {
    const int i = r * kSelectThreads + t;
    uint32_t key = 0xFFFFFFFFu;
    bool less = false, eq = false;
    if (i < n) {
      key = key_at(i);
      less = key < T;
      eq = key == T;
      const unsigned long long packed = ((unsigned long long)key << 32) | (uint32_t)i;
      my_min = packed < my_min ? packed : my_min;
    }
    const unsigned bl = __ballot_sync(0xFFFFFFFFu, less), be = __ballot_sync(0xFFFFFFFFu, eq);
    const unsigned lt_mask = (1u << lane) - 1u;
    if (lane == 0) { warp_less[warp] = __popc(bl); warp_eq[warp] = __popc(be); }
    __syncthreads();
    if (warp == 0) {
      // exclusive scan of the 32 per-warp counts; carry the running base across rounds
      const uint32_t cl = warp_less[lane], ce = warp_eq[lane];
      uint32_t il = cl, ie = ce;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t vl = __shfl_up_sync(0xFFFFFFFFu, il, d), ve = __shfl_up_sync(0xFFFFFFFFu, ie, d);
        if (lane >= d) { il += vl; ie += ve; }
      }
      const uint32_t bl0 = s_base_less, be0 = s_base_eq;
      __syncwarp();
      warp_less[lane] = bl0 + il - cl;
      warp_eq[lane] = be0 + ie - ce;
      if (lane == 31) { s_base_less = bl0 + il; s_base_eq = be0 + ie; }
    }
    __syncthreads();
    const uint32_t less_before = warp_less[warp] + __popc(bl & lt_mask);
    const uint32_t eq_before = warp_eq[warp] + __popc(be & lt_mask);
    if (less || (eq && eq_before < take_eq)) {
      const uint32_t pos = less_before + (eq_before < take_eq ? eq_before : take_eq);
      elite_idx[(long long)seg * k + pos] = i;
      if (elite_cost) elite_cost[(long long)seg * k + pos] = __ldg(c + i);
    }
    __syncthreads();  // warp_less / warp_eq are rewritten next round
  }

  // ---- block argmin (lowest index among equal minima) ----
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, my_min, d);
    my_min = o < my_min ? o : my_min;
  }
  if (lane == 0) warp_min[warp] = my_min;
  __syncthreads();
  if (warp == 0) {
    unsigned long long v = warp_min[lane];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, v, d);
      v = o < v ? o : v;
    }
    if (lane == 0 && n > 0) {
      const int idx = (int)(uint32_t)(v & 0xFFFFFFFFull);
      const float cmin = __ldg(c + idx);
      if (best) { best[seg].best_cost = cmin; best[seg].best_iteration = iteration; best[seg].best_index = idx; best[seg].reserved = 0; }
      if (best_ever) {
        BestEver b = best_ever[seg];
        if (b.iteration < 0 || cmin < b.cost) { b.cost = cmin; b.iteration = iteration; b.index = idx; best_ever[seg] = b; }
      }
    }
  }
}

// ---- refit ---------------------------------------------------------------------------
constexpr int kRefitThreads = 1024;

// grid = (H * G, E): one CTA per (step, 4-wide action group, env).  Each thread regenerates
// (or gathers) the 4 actions of its elites and accumulates shifted sums sum(d), sum(d^2) with
// d = a - c, c = the old mean of that (h, a) (the draws are centred there, so the shifted
// second moment does not cancel); a fixed shuffle tree + one shared-memory stage reduces them.
// mean = c + sum(d)/k, std = sqrt(sum(d^2)/k - (sum(d)/k)^2)   (population std, unbiased=False).
__global__ void __launch_bounds__(kRefitThreads)
refit_kernel(ActionSource src, Shape sh, int A, const int* __restrict__ elite_idx, int k,
             float* __restrict__ mu_new, float* __restrict__ sd_new) {
  __shared__ float red[kRefitThreads / 32][8];
  const int G = (A + 3) >> 2;
  const int h = blockIdx.x / G, g = blockIdx.x % G;
  const int env_l = blockIdx.y;
  const long long R = sh.rows();
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const long long ms = ((long long)env_l * sh.H + h) * A;
  const bool inject = src.mode == MBRL_SAMPLE_INJECT_ACTIONS || src.mode == MBRL_SAMPLE_INJECT_NOISE;
  const bool affine = src.mode == MBRL_SAMPLE_INJECT_NOISE || src.mode == MBRL_SAMPLE_GAUSSIAN;
  float mu_old[4], sd_old[4], s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ac = min(4 * g + j, A - 1);
    mu_old[j] = affine ? __ldg(src.mu + ms + ac) : 0.f;
    sd_old[j] = affine ? __ldg(src.sd + ms + ac) : 0.f;
  }
  const uint2 key = make_uint2(src.seed_lo, src.seed_hi);
  for (int e = t; e < k; e += kRefitThreads) {
    const int cand_l = __ldg(elite_idx + (long long)env_l * k + e);
    float z[4];
    if (inject) {
      const long long row = (long long)env_l * sh.N + cand_l;
      const float* p = src.buf + ((long long)h * R + row) * A;
#pragma unroll
      for (int j = 0; j < 4; ++j) z[j] = __ldg(p + min(4 * g + j, A - 1));
    } else {
      const uint4 r = philox4x32_10(make_uint4((uint32_t)(h * G + g), src.iteration,
                                               src.cand_offset + (uint32_t)cand_l,
                                               src.env_offset + (uint32_t)env_l), key);
      if (src.mode == MBRL_SAMPLE_GAUSSIAN) {
        const float4 q = box_muller4(r);
        z[0] = q.x; z[1] = q.y; z[2] = q.z; z[3] = q.w;
      } else {
        z[0] = u32_to_uniform(r.x); z[1] = u32_to_uniform(r.y);
        z[2] = u32_to_uniform(r.z); z[3] = u32_to_uniform(r.w);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = z[j];  // same arithmetic as the rollout kernels' samplers: bit-identical actions
      if (affine) v = clipf(__fadd_rn(mu_old[j], __fmul_rn(sd_old[j], v)), src.lo, src.hi);
      else if (src.mode == MBRL_SAMPLE_UNIFORM) v = __fadd_rn(src.lo, __fmul_rn(__fsub_rn(src.hi, src.lo), v));
      const float d = v - mu_old[j];
      s1[j] += d;
      s2[j] = fmaf(d, d, s2[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      s1[j] += __shfl_xor_sync(0xFFFFFFFFu, s1[j], d);
      s2[j] += __shfl_xor_sync(0xFFFFFFFFu, s2[j], d);
    }
    if (lane == 0) { red[warp][j] = s1[j]; red[warp][4 + j] = s2[j]; }
  }
  __syncthreads();
  if (t < 4) {
    float a1 = 0.f, a2 = 0.f;
    for (int w = 0; w < kRefitThreads / 32; ++w) { a1 += red[w][t]; a2 += red[w][4 + t]; }
    const int a = 4 * g + t;
    const float c = t == 0 ? mu_old[0] : t == 1 ? mu_old[1] : t == 2 ? mu_old[2] : mu_old[3];
    if (a < A) {
      const float inv = 1.0f / (float)k, m1 = a1 * inv;
      mu_new[ms + a] = c + m1;
      sd_new[ms + a] = __fsqrt_rn(fmaxf(fmaf(-m1, m1, a2 * inv), 0.f));
    }
  }
}

// mu/sd initialisation: (lo+hi)/2 and (hi-lo)/2 per (env, h, a); best-ever reset.
__global__ void init_plan_kernel(float* __restrict__ mu, float* __restrict__ sd, long long n,
                                 float lo, float hi, BestEver* __restrict__ best_ever, int E) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (mu && i < n) { mu[i] = 0.5f * (lo + hi); sd[i] = 0.5f * (hi - lo); }
  if (best_ever && i < E) best_ever[i] = BestEver{0.f, -1, -1, 0};
}

}  // namespace mbrl
