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

// One CTA per segment (environment).  4 radix passes of 8 bits find the k-th smallest key
// T and the number of keys strictly below it; a final index-ordered compaction emits every
// key < T plus the first (k - count_less) keys == T.  Warp-shuffle/ballot scans, shared
// memory only for the 256-bin histogram and per-warp carries.
//   best (nullable):      (min cost, ., argmin) of this launch per segment
//   best_ever (nullable): updated when this launch's minimum is strictly smaller
//                         (earlier iteration wins ties)
__global__ void __launch_bounds__(kSelectThreads)
topk_select_kernel(const float* __restrict__ costs, int n, int k, int* __restrict__ elite_idx,
                   float* __restrict__ elite_cost, MbrlPlanInfo* __restrict__ best,
                   BestEver* __restrict__ best_ever, int iteration) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_prefix, s_mask, s_remaining;
  __shared__ uint32_t warp_less[32], warp_eq[32];
  __shared__ unsigned long long warp_min[32];
  __shared__ uint32_t s_base_less, s_base_eq;

  const int seg = blockIdx.x;
  const float* c = costs + (long long)seg * n;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

  if (t == 0) { s_prefix = 0; s_mask = 0; s_remaining = (uint32_t)k; }

  // ---- radix select, MSB first ----
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (t < 256) hist[t] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix, mask = s_mask;
    for (int base = 0; base < n; base += kSelectThreads) {  // warp-uniform trip count
      const int i = base + t;
      uint32_t key = 0, bin = 0;
      bool in = false;
      if (i < n) {
        key = cost_key(__ldg(c + i));
        in = (key & mask) == prefix;
        bin = (key >> shift) & 0xFFu;
      }
      // warp-aggregated histogram update: one shared-memory atomic per distinct bin
      const unsigned active = __ballot_sync(0xFFFFFFFFu, in);
      if (in) {
        const unsigned peers = __match_any_sync(active, bin);
        if (lane == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
      }
      __syncwarp();
    }
    __syncthreads();
    if (warp == 0) {
      // 8 bins per lane, exclusive scan across the warp, locate the bin holding the
      // remaining-th smallest
      uint32_t local[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { local[j] = hist[lane * 8 + j]; sum += local[j]; }
      uint32_t incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += v;
      }
      uint32_t run = incl - sum;
      const uint32_t rem = s_remaining;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (rem > run && rem <= run + local[j]) {
          s_prefix = prefix | ((uint32_t)(lane * 8 + j) << shift);
          s_mask = mask | (0xFFu << shift);
          s_remaining = rem - run;  // rank inside the chosen bin
        }
        run += local[j];
      }
    }
    __syncthreads();
  }
  const uint32_t T = s_prefix;         // k-th smallest key
  const uint32_t take_eq = s_remaining;  // how many keys == T belong to the elite set
  if (t == 0) { s_base_less = 0; s_base_eq = 0; }
  __syncthreads();

  // ---- index-ordered compaction + argmin ----
  unsigned long long my_min = ~0ull;
  const int rounds = (n + kSelectThreads - 1) / kSelectThreads;
  for (int r = 0; r < rounds; ++r) {
    const int i = r * kSelectThreads + t;
    float cv = 0.f;
    uint32_t key = 0xFFFFFFFFu;
    bool less = false, eq = false;
    if (i < n) {
      cv = __ldg(c + i);
      key = cost_key(cv);
      less = key < T;
      eq = key == T;
      const unsigned long long packed = ((unsigned long long)key << 32) | (uint32_t)i;
      my_min = packed < my_min ? packed : my_min;
    }
    const unsigned bl = __ballot_sync(0xFFFFFFFFu, less), be = __ballot_sync(0xFFFFFFFFu, eq);
    const unsigned lt_mask = (1u << lane) - 1u;
    if (lane == 0) { warp_less[warp] = __popc(bl); warp_eq[warp] = __popc(be); }
    __syncthreads();
    uint32_t less_before = s_base_less, eq_before = s_base_eq;
    for (int w = 0; w < warp; ++w) { less_before += warp_less[w]; eq_before += warp_eq[w]; }
    less_before += __popc(bl & lt_mask);
    eq_before += __popc(be & lt_mask);
    const bool sel = less || (eq && eq_before < take_eq);
    if (sel) {
      const uint32_t pos = less_before + (eq_before < take_eq ? eq_before : take_eq);
      elite_idx[(long long)seg * k + pos] = i;
      if (elite_cost) elite_cost[(long long)seg * k + pos] = cv;
    }
    __syncthreads();
    if (t == kSelectThreads - 1) {
      s_base_less = less_before + (less ? 1u : 0u);
      s_base_eq = eq_before + (eq ? 1u : 0u);
    }
    __syncthreads();
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
// Welford accumulator with Chan's pairwise merge: single pass, deterministic tree order.
struct Moments {
  float n, mean, m2;
};
__device__ __forceinline__ void moments_push(Moments& m, float x) {
  m.n += 1.0f;
  const float d = x - m.mean;
  m.mean += d / m.n;
  m.m2 += d * (x - m.mean);
}
__device__ __forceinline__ Moments moments_merge(const Moments& a, const Moments& b) {
  if (b.n == 0.f) return a;
  if (a.n == 0.f) return b;
  Moments r;
  r.n = a.n + b.n;
  const float d = b.mean - a.mean;
  r.mean = a.mean + d * (b.n / r.n);
  r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / r.n);
  return r;
}
__device__ __forceinline__ Moments moments_shfl_xor(const Moments& m, int d) {
  Moments o;
  o.n = __shfl_xor_sync(0xFFFFFFFFu, m.n, d);
  o.mean = __shfl_xor_sync(0xFFFFFFFFu, m.mean, d);
  o.m2 = __shfl_xor_sync(0xFFFFFFFFu, m.m2, d);
  return o;
}

constexpr int kRefitThreads = 256;

// grid = (H * G, E): one CTA per (step, 4-wide action group, env).  Each thread regenerates
// (or gathers) the 4 actions of its elites, then a shuffle tree merges the moments.
// mean = sum/k, std = sqrt(sum((a-mean)^2)/k)  (population std, unbiased=False).
__global__ void __launch_bounds__(kRefitThreads)
refit_kernel(ActionSource src, Shape sh, int A, const int* __restrict__ elite_idx, int k,
             float* __restrict__ mu_new, float* __restrict__ sd_new) {
  __shared__ Moments red[kRefitThreads / 32][4];
  const int G = (A + 3) >> 2;
  const int h = blockIdx.x / G, g = blockIdx.x % G;
  const int env_l = blockIdx.y;
  const long long R = sh.rows();
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  Moments acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = Moments{0.f, 0.f, 0.f};

  const long long ms = ((long long)env_l * sh.H + h) * A;
  const uint2 key = make_uint2(src.seed_lo, src.seed_hi);
  for (int e = t; e < k; e += kRefitThreads) {
    const int cand_l = __ldg(elite_idx + (long long)env_l * k + e);
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (src.mode == MBRL_SAMPLE_INJECT_ACTIONS || src.mode == MBRL_SAMPLE_INJECT_NOISE) {
      const long long row = (long long)env_l * sh.N + cand_l;
      const float* p = src.buf + ((long long)h * R + row) * A;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int a = 4 * g + j;
        if (a < A) {
          const float x = __ldg(p + a);
          v[j] = src.mode == MBRL_SAMPLE_INJECT_ACTIONS
                     ? x
                     : clipf(__fadd_rn(__ldg(src.mu + ms + a), __fmul_rn(__ldg(src.sd + ms + a), x)), src.lo, src.hi);
        }
      }
    } else {
      const uint4 r = philox4x32_10(make_uint4((uint32_t)(h * G + g), src.iteration,
                                               src.cand_offset + (uint32_t)cand_l,
                                               src.env_offset + (uint32_t)env_l), key);
      float z[4];
      if (src.mode == MBRL_SAMPLE_GAUSSIAN) {
        const float4 q = box_muller4(r);
        z[0] = q.x; z[1] = q.y; z[2] = q.z; z[3] = q.w;
      } else {
        z[0] = u32_to_uniform(r.x); z[1] = u32_to_uniform(r.y);
        z[2] = u32_to_uniform(r.z); z[3] = u32_to_uniform(r.w);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int a = 4 * g + j;
        if (a < A) {
          v[j] = src.mode == MBRL_SAMPLE_GAUSSIAN
                     ? clipf(__fadd_rn(__ldg(src.mu + ms + a), __fmul_rn(__ldg(src.sd + ms + a), z[j])), src.lo, src.hi)
                     : __fadd_rn(src.lo, __fmul_rn(__fsub_rn(src.hi, src.lo), z[j]));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) moments_push(acc[j], v[j]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc[j] = moments_merge(acc[j], moments_shfl_xor(acc[j], d));
    if (lane == 0) red[warp][j] = acc[j];
  }
  __syncthreads();
  if (t < 4) {
    Moments m = red[0][t];
    for (int w = 1; w < kRefitThreads / 32; ++w) m = moments_merge(m, red[w][t]);
    const int a = 4 * g + t;
    if (a < A) {
      mu_new[ms + a] = m.mean;
      sd_new[ms + a] = __fsqrt_rn(m.m2 / m.n);
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
