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
#ifdef MBRL_TOPK_PROFILE
__device__ long long g_topk_stamps[32];
#define TOPK_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_topk_stamps[i] = clock64(); } while (0)
#else
#define TOPK_STAMP(i) do { } while (0)
#endif

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
// range [lo, hi] (power-of-two width; the candidates spread over the bins instead of piling onto
// the few leading bit patterns that the costs of one population share), a block scan locates the
// bucket holding the k-th key, and the range shrinks >= 512x; it ends when the width is 1.
// A final index-ordered compaction (ballot + warp-shuffle scans) emits every key < T plus the
// first `take_eq` keys == T.
//   best (nullable):      (min cost, ., argmin) of this launch per segment
//   best_ever (nullable): updated when this launch's minimum is strictly smaller
//                         (earlier iteration wins ties)
// STAGED=false re-reads the costs from global memory (segments too long for shared memory).
constexpr int kSelectStageMax = 49152;  // keys staged in shared memory: 192 KB
constexpr int kSelectBins = 1024;

// inclusive warp scan
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, v, d);
    if (lane >= d) v += o;
  }
  return v;
}

template <bool STAGED>
__global__ void __launch_bounds__(kSelectThreads)
topk_select_kernel(const float* __restrict__ costs, int n, int k, int* __restrict__ elite_idx,
                   float* __restrict__ elite_cost, MbrlPlanInfo* __restrict__ best,
                   BestEver* __restrict__ best_ever, int iteration) {
  extern __shared__ __align__(16) uint32_t sel_smem[];
  uint32_t* keys = sel_smem;  // [round4(n)] when STAGED
  __shared__ uint32_t hist[kSelectBins];
  __shared__ uint32_t wtot[2][4][32];  // per-warp totals (double buffered; 4 trips per pass)
  __shared__ uint32_t wtot2[2][4][32];
  __shared__ uint32_t s_sel[2];        // winning bin, remaining rank
  __shared__ unsigned long long warp_min[32];

  const int seg = blockIdx.x;
  const float* c = costs + (long long)seg * n;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int n4 = (n + 3) & ~3;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(c) & 15) == 0);
  pdl_trigger();
  pdl_wait();  // the costs come from the preceding rollout / unpack kernel
  TOPK_STAMP(0);

  // four keys of indices i4..i4+3 (i4 multiple of 4); out-of-range -> 0xFFFFFFFF (masked by index)
  auto load4 = [&](int i4, uint32_t (&kk)[4]) {
    if (vec_ok && i4 + 3 < n) {
      const float4 q = dep_load(reinterpret_cast<const float4*>(c + i4));
      kk[0] = cost_key(q.x); kk[1] = cost_key(q.y); kk[2] = cost_key(q.z); kk[3] = cost_key(q.w);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) kk[j] = i4 + j < n ? cost_key(dep_load(c + i4 + j)) : 0xFFFFFFFFu;
    }
  };
  auto key4_at = [&](int i4, uint32_t (&kk)[4]) {
    if (STAGED) {
      const uint4 q = *reinterpret_cast<const uint4*>(keys + i4);
      kk[0] = q.x; kk[1] = q.y; kk[2] = q.z; kk[3] = q.w;
    } else {
      load4(i4, kk);
    }
  };

  // ---- stage + min/max (16-byte loads, all in flight together) ----
  uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
#pragma unroll 4
  for (int i4 = 4 * t; i4 < n4; i4 += 4 * kSelectThreads) {
    uint32_t kk[4];
    load4(i4, kk);
    if (STAGED) *reinterpret_cast<uint4*>(keys + i4) = make_uint4(kk[0], kk[1], kk[2], kk[3]);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i4 + j < n) { kmin = min(kmin, kk[j]); kmax = max(kmax, kk[j]); }
  }
  hist[t] = 0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xFFFFFFFFu, kmin, d));
    kmax = max(kmax, __shfl_xor_sync(0xFFFFFFFFu, kmax, d));
  }
  if (lane == 0) { wtot[0][0][warp] = kmin; wtot[0][1][warp] = kmax; }
  __syncthreads();
  kmin = wtot[0][0][lane]; kmax = wtot[0][1][lane];  // every warp finishes the reduction itself
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xFFFFFFFFu, kmin, d));
    kmax = max(kmax, __shfl_xor_sync(0xFFFFFFFFu, kmax, d));
  }
  uint32_t lo = kmin, hi = kmax, rem = (uint32_t)k;
  TOPK_STAMP(1);

  // ---- adaptive radix select: power-of-two bucket width, <= 1024 buckets over [lo, hi] ----
  for (int round = 0; round < 6; ++round) {
    const uint32_t span = hi - lo;                              // in-range test: key - lo <= span
    const int shift = span < (uint32_t)kSelectBins ? 0 : 32 - __clz(span) - 10;  // span >> shift < 1024
    TOPK_STAMP(2 + 3 * round);
    for (int i4 = 4 * t; i4 < n4; i4 += 4 * kSelectThreads) {
      uint32_t kk[4];
      key4_at(i4, kk);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t d = kk[j] - lo;
        if (d <= span && i4 + j < n) atomicAdd(&hist[d >> shift], 1u);
      }
    }
    __syncthreads();
    TOPK_STAMP(3 + 3 * round);
    // block-wide scan of the 1024 bins: one bin per thread, warp totals through shared memory,
    // every warp scans the 32 totals itself (no single-warp phase)
    const uint32_t mine = hist[t];
    hist[t] = 0;  // ready for the next round
    const uint32_t incl_w = warp_incl_scan(mine, lane);
    const int buf = round & 1;
    if (lane == 31) wtot[buf][0][warp] = incl_w;
    __syncthreads();
    const uint32_t tot = wtot[buf][0][lane];
    const uint32_t tot_incl = warp_incl_scan(tot, lane);
    const uint32_t base = __shfl_sync(0xFFFFFFFFu, tot_incl - tot, warp);
    const uint32_t incl = base + incl_w, excl = incl - mine;
    if (rem > excl && rem <= incl) { s_sel[0] = (uint32_t)t; s_sel[1] = rem - excl; }  // exactly one bin
    __syncthreads();
    const uint32_t bin = s_sel[0];
    rem = s_sel[1];
    const uint32_t nspan = min(span - (bin << shift), (1u << shift) - 1u);
    lo = lo + (bin << shift);
    hi = lo + nspan;
    TOPK_STAMP(4 + 3 * round);
    if (shift == 0) break;  // uniform: buckets were single key values
  }
  const uint32_t T = lo;          // k-th smallest key
  const uint32_t take_eq = rem;   // how many keys == T belong to the elite set
  TOPK_STAMP(20);

  // ---- index-ordered compaction + argmin ----
  // A pass covers 4 trips of 4096 indices (thread t owns indices trip*4096 + 4t .. +3); the per-warp
  // counts of all 4 trips meet in shared memory once, every warp scans them itself, and the
  // running base lives in registers: one barrier per 16384 keys.
  unsigned long long my_min = ~0ull;
  uint32_t base_less = 0, base_eq = 0;
  int pass = 0;
  for (int p0 = 0; p0 < n4; p0 += 16 * kSelectThreads, ++pass) {
    uint32_t kk[4][4], il[4], ie[4], nl[4], ne[4];
#pragma unroll
    for (int tr = 0; tr < 4; ++tr) {
      const int i4 = p0 + tr * 4 * kSelectThreads + 4 * t;
      nl[tr] = 0; ne[tr] = 0;
      if (i4 < n4) {
        key4_at(i4, kk[tr]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (i4 + j < n) {
            nl[tr] += kk[tr][j] < T;
            ne[tr] += kk[tr][j] == T;
            const unsigned long long packed = ((unsigned long long)kk[tr][j] << 32) | (uint32_t)(i4 + j);
            my_min = packed < my_min ? packed : my_min;
          }
        }
      }
      il[tr] = warp_incl_scan(nl[tr], lane);
      ie[tr] = warp_incl_scan(ne[tr], lane);
      if (lane == 31) { wtot[pass & 1][tr][warp] = il[tr]; wtot2[pass & 1][tr][warp] = ie[tr]; }
    }
    __syncthreads();
#pragma unroll
    for (int tr = 0; tr < 4; ++tr) {
      const int i4 = p0 + tr * 4 * kSelectThreads + 4 * t;
      const uint32_t tl = wtot[pass & 1][tr][lane], te = wtot2[pass & 1][tr][lane];
      const uint32_t sl = warp_incl_scan(tl, lane), se = warp_incl_scan(te, lane);
      uint32_t less_before = base_less + __shfl_sync(0xFFFFFFFFu, sl - tl, warp) + il[tr] - nl[tr];
      uint32_t eq_before = base_eq + __shfl_sync(0xFFFFFFFFu, se - te, warp) + ie[tr] - ne[tr];
      base_less += __shfl_sync(0xFFFFFFFFu, sl, 31);
      base_eq += __shfl_sync(0xFFFFFFFFu, se, 31);
      if (i4 < n4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = i4 + j;
          if (i < n) {
            const bool less = kk[tr][j] < T, eq = kk[tr][j] == T;
            if (less || (eq && eq_before < take_eq)) {
              const uint32_t pos = less_before + (eq_before < take_eq ? eq_before : take_eq);
              elite_idx[(long long)seg * k + pos] = i;
              if (elite_cost) elite_cost[(long long)seg * k + pos] = dep_load(c + i);
            }
            less_before += less;
            eq_before += eq;
          }
        }
      }
    }
  }
  TOPK_STAMP(21);

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
      const float cmin = dep_load(c + idx);
      if (best) { best[seg].best_cost = cmin; best[seg].best_iteration = iteration; best[seg].best_index = idx; best[seg].reserved = 0; }
      if (best_ever) {
        BestEver b = best_ever[seg];
        if (b.iteration < 0 || cmin < b.cost) { b.cost = cmin; b.iteration = iteration; b.index = idx; best_ever[seg] = b; }
      }
    }
  }
  TOPK_STAMP(22);
}

// ---- refit ---------------------------------------------------------------------------
constexpr int kRefitThreads = 1024;
constexpr int kRefitChunk = 2 * kRefitThreads;  // elites per CTA
// CTA width for k elites: one elite per thread up to 1024 (a function of k alone, so the summation
// order -- and with it the bit pattern of the refit -- does not depend on how the population is
// sharded).  Small elite sets (k = 204 at the cfg-5 shard) would otherwise pay for 1024-thread CTAs
// that are 80 % idle: 38 400 of them took 0.8 ms per iteration.
inline int refit_threads(int k) { return k >= kRefitThreads ? kRefitThreads : (k < 1 ? 1 : k + 31) / 32 * 32; }

// grid = (H * G, E, chunks): one CTA per (step, 4-wide action group, env, chunk of 2048 elites).
// Each thread regenerates (or gathers) the 4 actions of its elites and accumulates shifted sums
// sum(d), sum(d^2) with d = a - c, c = the old mean of that (h, a) (the draws are centred there, so
// the shifted second moment does not cancel); a fixed shuffle tree + one shared-memory stage
// reduces them.  mean = c + sum(d)/k, std = sqrt(sum(d^2)/k - (sum(d)/k)^2)   (population std).
// k > 2048 (large or population-sharded elite sets): the chunk CTAs run in parallel on otherwise
// idle SMs, park their partial sums, and the last one to arrive adds them IN CHUNK ORDER -- the
// result depends only on the elite list (ascending index), never on timing or on the sharding,
// which is what keeps every rank's refit bit-identical to the unsharded one.
__global__ void __launch_bounds__(kRefitThreads)
refit_kernel(ActionSource src, Shape sh, int A, const int* __restrict__ elite_idx, int k,
             float* __restrict__ mu_new, float* __restrict__ sd_new, float* __restrict__ part,
             unsigned int* __restrict__ arrive) {
  __shared__ float red[kRefitThreads / 32][8];
  __shared__ bool s_last;
  const int G = (A + 3) >> 2;
  const int h = blockIdx.x / G, g = blockIdx.x % G;
  const int env_l = blockIdx.y;
  const int chunk = blockIdx.z, nchunks = gridDim.z;
  const int e_end = min(k, (chunk + 1) * kRefitChunk);
  const long long R = sh.rows();
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nthreads = blockDim.x;
  const long long ms = ((long long)env_l * sh.H + h) * A;
  const bool inject = src.mode == MBRL_SAMPLE_INJECT_ACTIONS || src.mode == MBRL_SAMPLE_INJECT_NOISE;
  const bool affine = src.mode == MBRL_SAMPLE_INJECT_NOISE || src.mode == MBRL_SAMPLE_GAUSSIAN;
  float mu_old[4], sd_old[4], s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  pdl_trigger();
  pdl_wait();  // elite indices (top-k / remap) and the old mean/std (previous refit)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ac = min(4 * g + j, A - 1);
    mu_old[j] = affine ? dep_load(src.mu + ms + ac) : 0.f;
    sd_old[j] = affine ? dep_load(src.sd + ms + ac) : 0.f;
  }
  const uint2 key = make_uint2(src.seed_lo, src.seed_hi);
  for (int e = chunk * kRefitChunk + t; e < e_end; e += nthreads) {
    const int cand_l = dep_load(elite_idx + (long long)env_l * k + e);
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
  float a1 = 0.f, a2 = 0.f;
  if (t < 4)
    for (int w = 0; w < nthreads / 32; ++w) { a1 += red[w][t]; a2 += red[w][4 + t]; }
  if (nchunks > 1) {
    const long long slot = (long long)env_l * gridDim.x + blockIdx.x;
    float* mine = part + (slot * nchunks + chunk) * 8;
    if (t < 4) { mine[t] = a1; mine[4 + t] = a2; }
    __threadfence();
    __syncthreads();
    if (t == 0) {
      const unsigned int n = atomicAdd(arrive + slot, 1u);
      s_last = n == (unsigned int)nchunks - 1;
      if (s_last) arrive[slot] = 0;  // ready for the next launch (stream order)
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (t < 4) {
      a1 = 0.f; a2 = 0.f;
      for (int c = 0; c < nchunks; ++c) {
        a1 += __ldcg(part + (slot * nchunks + c) * 8 + t);
        a2 += __ldcg(part + (slot * nchunks + c) * 8 + 4 + t);
      }
    }
  }
  if (t < 4) {
    const int a = 4 * g + t;
    const float c = t == 0 ? mu_old[0] : t == 1 ? mu_old[1] : t == 2 ? mu_old[2] : mu_old[3];
    if (a < A) {
      const float inv = 1.0f / (float)k, m1 = a1 * inv;
      mu_new[ms + a] = c + m1;
      sd_new[ms + a] = __fsqrt_rn(fmaxf(fmaf(-m1, m1, a2 * inv), 0.f));
    }
  }
}

// ---- population-sharded elite merge (see mbrl_comm_init) ---------------------------------------
// send buffer of one rank: [k_l costs (bits) | k_l global indices]
__global__ void pack_elites_kernel(const float* __restrict__ elite_cost, const int* __restrict__ elite_idx,
                                   int k_l, int idx_offset, uint32_t* __restrict__ send) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (i < k_l) {
    send[i] = __float_as_uint(dep_load(elite_cost + i));
    send[k_l + i] = (uint32_t)(dep_load(elite_idx + i) + idx_offset);
  }
}
// ---- peer-memory (NVLink P2P) elite exchange ---------------------------------------------------
// Layout of every rank's exported buffer: [2 parities][world][2*slot] uint32 data, then
// [world] uint32 sequence flags (one per source rank).  slot = k_l capacity.
struct P2pPeers {
  uint32_t* base[64];  // peer r's exported buffer (own rank: the local pointer)
};
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// One launch per iteration: writes this rank's k_l (cost bits | global index) pairs into slot
// [parity][rank] of EVERY rank's buffer (peer stores over NVLink), then -- after a system-scope
// fence and a grid-wide arrival count -- the last block publishes `seq` in flag[rank] of every peer.
__global__ void p2p_scatter_kernel(const float* __restrict__ elite_cost, const int* __restrict__ elite_idx,
                                   int k_l, int idx_offset, P2pPeers peers, int rank, int world, int slot,
                                   int parity, uint32_t seq, unsigned int* __restrict__ arrive_counter) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (i < k_l) {
    const uint32_t c = __float_as_uint(dep_load(elite_cost + i));
    const uint32_t g = (uint32_t)(dep_load(elite_idx + i) + idx_offset);
    const size_t off = ((size_t)parity * world + rank) * 2 * slot;
    for (int r = 0; r < world; ++r) {
      uint32_t* dst = peers.base[r] + off;
      dst[i] = c;
      dst[k_l + i] = g;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(arrive_counter, 1u) + 1u;
    if (done == gridDim.x) {
      *arrive_counter = 0;  // ready for the next launch (stream order)
      __threadfence_system();
      const size_t flags = (size_t)2 * world * 2 * slot;
      for (int r = 0; r < world; ++r) st_release_sys(peers.base[r] + flags + rank, seq);
    }
  }
}
// Consumer: acquire every rank's flag (>= seq), then unpack [parity][r][..] into contiguous costs /
// global indices in rank order.  A rank that never shows up within `timeout_ns` (wall clock, default
// 120 s, MBRL_P2P_TIMEOUT_S) trips the timeout: *error = 1 and the gathered arrays are filled with
// (+inf, -1) sentinels instead of the previous iteration's data, so that whatever runs next is
// deterministic garbage that the plan reports (info.reserved bit 1) rather than a plausible wrong plan.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__global__ void p2p_wait_unpack_kernel(const uint32_t* __restrict__ local, int world, int slot, int k_l, int parity,
                                       uint32_t seq, float* __restrict__ gcost, int* __restrict__ gidx,
                                       int* __restrict__ error, unsigned long long timeout_ns) {
  __shared__ int s_ok;
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (threadIdx.x < world) {
    const uint32_t* flag = local + (size_t)2 * world * 2 * slot + threadIdx.x;
    const unsigned long long t0 = globaltimer_ns();
    unsigned int spins = 0;
    while ((int)(ld_acquire_sys(flag) - seq) < 0) {
      if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > timeout_ns) { s_ok = 0; break; }
    }
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (!s_ok) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *error = 1;
    if (i < world * k_l) { gcost[i] = __int_as_float(0x7f800000); gidx[i] = -1; }
    return;
  }
  if (i < world * k_l) {
    const int r = i / k_l, j = i - r * k_l;
    const uint32_t* src = local + ((size_t)parity * world + r) * 2 * slot;
    gcost[i] = __uint_as_float(__ldcg(src + j));
    gidx[i] = (int)__ldcg(src + k_l + j);
  }
}

// End of a sharded plan: info.reserved = (reduced-gather-not-provably-exact) | (exchange timed out) << 1;
// the truncation flag is reset for the next plan whether or not the caller passed an info buffer.
__global__ void shard_flags_kernel(MbrlPlanInfo* __restrict__ info, int* __restrict__ trunc, const int* __restrict__ p2p_error) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const int t = dep_load(trunc) != 0, e = p2p_error ? (dep_load(p2p_error) != 0) : 0;
    if (info) info[0].reserved = t | (e << 1);
    *trunc = 0;
  }
}

// gathered [world][2*k_l] -> contiguous costs / global indices in rank order (== ascending global
// index, so "ties -> lower position" in the merge is "ties -> lower global index")
__global__ void unpack_gathered_kernel(const uint32_t* __restrict__ recv, int world, int k_l,
                                       float* __restrict__ gcost, int* __restrict__ gidx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (i < world * k_l) {
    const int r = i / k_l, j = i - r * k_l;
    gcost[i] = __uint_as_float(dep_load(recv + (long long)r * 2 * k_l + j));
    gidx[i] = (int)dep_load(recv + (long long)r * 2 * k_l + k_l + j);
  }
}
// positions in the gathered list -> global candidate indices; best-ever bookkeeping in global indices
// Truncation check: each rank sent only its k_s cheapest (k_s < the worst-case min(k, N)); the merge
// is exact unless ALL k_s candidates of some rank were selected (then cheaper-than-threshold
// candidates of that rank may have been left out) -> *trunc_flag = 1, the caller redoes the plan
// with full-size gathers.  pos is ascending, so per-rank counts are two binary searches.
__global__ void remap_elites_kernel(const int* __restrict__ pos, const int* __restrict__ gidx, int k,
                                    int* __restrict__ elite_global, const MbrlPlanInfo* __restrict__ best_now,
                                    BestEver* __restrict__ best_ever, int iteration, int world, int k_s,
                                    int k_full, int* __restrict__ trunc_flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (i < k) elite_global[i] = dep_load(gidx + dep_load(pos + i));
  if (i < world && k_s < k_full) {
    auto lower_bound = [&](int v) {
      int lo = 0, hi = k;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (dep_load(pos + mid) < v) lo = mid + 1; else hi = mid; }
      return lo;
    };
    if (lower_bound((i + 1) * k_s) - lower_bound(i * k_s) == k_s) atomicOr(trunc_flag, 1);
  }
  if (i == 0) {
    const float cmin = dep_load(&best_now->best_cost);
    BestEver b = *best_ever;
    if (b.iteration < 0 || cmin < b.cost) {
      b.cost = cmin; b.iteration = iteration; b.index = dep_load(gidx + dep_load(&best_now->best_index));
      *best_ever = b;
    }
  }
}

// mu/sd initialisation: (lo+hi)/2 and (hi-lo)/2 per (env, h, a); best-ever reset.
__global__ void init_plan_kernel(float* __restrict__ mu, float* __restrict__ sd, long long n,
                                 float lo, float hi, BestEver* __restrict__ best_ever, int E) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (mu && i < n) { mu[i] = 0.5f * (lo + hi); sd[i] = 0.5f * (hi - lo); }
  if (best_ever && i < E) best_ever[i] = BestEver{0.f, -1, -1, 0};
}

// Warm start resident on the device (MBRL_WARM_USE): the new mean is the previous plan's final mean
// shifted by one step with the last step repeated (the MPC receding-horizon shift of
// MPCPolicy's hand-over, src/mbrl/agents.py:41-47), std = `std`; best-ever reset.
__global__ void warm_init_kernel(float* __restrict__ mu, float* __restrict__ sd, const float* __restrict__ last_mu,
                                 int E, int H, int A, float std, float lo, float hi, BestEver* __restrict__ best_ever) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  pdl_trigger();
  pdl_wait();
  const long long n = (long long)E * H * A;
  if (i < n) {
    const int a = (int)(i % A), h = (int)((i / A) % H);
    const long long e = i / ((long long)A * H);
    const int hs = h + 1 < H ? h + 1 : H - 1;
    mu[i] = clipf(dep_load(last_mu + (e * H + hs) * A + a), lo, hi);
    sd[i] = std;
  }
  if (best_ever && i < E) best_ever[i] = BestEver{0.f, -1, -1, 0};
}

}  // namespace mbrl
