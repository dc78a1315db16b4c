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

// ---- peer-memory (NVLink P2P) elite exchange ---------------------------------------------------
// (mbrl_p2p_export / mbrl_p2p_attach).  Every rank exports one buffer made of PACKETS: 64-bit words
// {value (32 bits), sequence tag (32 bits)} written with one scalar store each, so a reader that
// finds this iteration's tag holds this iteration's value -- no flags and no system-scope fences;
// data is consumed as it lands.  Per iteration parity (double buffering):
//   cost packets  [world][k_l]   rank r's k_l cheapest costs in ascending candidate order, packed
//                                contiguously (k_l <= slot): the gathered costs are one array of
//                                world*k_l entries in ascending GLOBAL index order
//   index packets [k_l]          the global indices of THIS rank's entries (written locally)
//   rank packets  [world][3]     rank r's local threshold key, minimum cost, argmin (global index)
//   refit packets [world][pslots][8]  partial sums; pslots = H * ceil(A/4) (step, action group) slots
// The parity half that a sequence number selects was last read two iterations earlier (a writer
// cannot be two iterations ahead of a reader: its own merge needs every rank's packets of the
// iteration in between), so a reader sees this iteration's tag or an older one, never a newer one.
struct P2pPeers {
  uint32_t* base[64];  // peer r's exported buffer (own rank: the local pointer)
};
// one naturally aligned 64-bit SCALAR access (vector accesses carry no single-copy atomicity in the
// PTX memory model): a {value, sequence tag} packet is never seen torn
__device__ __forceinline__ void st_packet(uint32_t* p, uint32_t v, uint32_t tag) {
  const unsigned long long w = (unsigned long long)v | ((unsigned long long)tag << 32);
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ uint2 ld_packet(const uint32_t* p) {
  unsigned long long w;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return make_uint2((uint32_t)w, (uint32_t)(w >> 32));
}
__host__ __device__ inline size_t p2p_idx_off(int world, int slot) { return 2 * (size_t)world * slot; }
__host__ __device__ inline size_t p2p_rank_off(int world, int slot) { return p2p_idx_off(world, slot) + 2 * (size_t)slot; }
__host__ __device__ inline size_t p2p_part_off(int world, int slot) { return p2p_rank_off(world, slot) + 6 * (size_t)world; }
__host__ __device__ inline size_t p2p_parity_words(int world, int slot, int pslots) {
  return (p2p_part_off(world, slot) + 16 * (size_t)world * pslots + 3) & ~(size_t)3;  // 16-byte multiple
}
__host__ __device__ inline size_t p2p_total_words(int world, int slot, int pslots) { return 2 * p2p_parity_words(world, slot, pslots); }
// Spin until the packet carries `seq`; false (and the stale value) once timeout_ns of wall clock have passed since t0.
__device__ __forceinline__ bool wait_packet(const uint32_t* p, uint32_t seq, unsigned long long t0,
                                            unsigned long long timeout_ns, uint32_t& value) {
  uint2 pkt = ld_packet(p);
  unsigned int spins = 0;
  while (pkt.y != seq) {
    if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > timeout_ns) { value = pkt.x; return false; }
    pkt = ld_packet(p);
  }
  value = pkt.x;
  return true;
}

// Sharded roles of the top-k kernel (template parameter MODE):
//   kSelPlain    the ordinary segmented top-k
//   kSelScatter  local top-k_l whose compaction stores cost packets straight into every rank's
//                exported buffer over NVLink (+ its own index packets, threshold and minimum)
//   kSelMerge    stages the world*k_l gathered cost packets as they arrive, finds the global top-k
//                threshold, emits the GLOBAL indices of THIS rank's elites only (+ their count: the
//                refit is distributed), keeps the best-ever record in global indices and checks
//                that the reduced-size gather was exact
enum { kSelPlain = 0, kSelScatter = 1, kSelMerge = 2 };
__device__ __forceinline__ bool k_per_rank_lt(int n, int world, int k_full) { return n / world < k_full; }
struct SelShard {
  P2pPeers peers;          // kSelScatter: every rank's buffer (own rank: the local pointer)
  const uint32_t* local;   // kSelMerge: this rank's buffer
  int rank, world, slot, pslots, parity;
  int idx_offset;          // kSelScatter: global index of local candidate 0
  int k_full;              // kSelMerge: min(k, N) -- a gather of k_full per rank is exact by construction
  uint32_t seq;
  int* trunc;              // kSelMerge: set when the reduced gather cannot be proven exact
  int* error;              // kSelMerge: set when a rank's packets never arrived
  int* own_count;          // kSelMerge: number of this rank's elites
  unsigned long long timeout_ns;
  long long* stamps;       // diagnostic (MBRL_SHARD_TIMELINE): globaltimer stamps of this iteration, or null
};
#define SHARD_STAMP(ptr, i) do { if ((ptr) && threadIdx.x == 0 && blockIdx.x == 0) (ptr)[i] = (long long)globaltimer_ns(); } while (0)

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
// histograms the keys that are still in range into 4096 equal-width buckets of the CURRENT key
// range [lo, hi] (power-of-two width; the candidates spread over the bins instead of piling onto
// the few leading bit patterns that the costs of one population share), a block scan locates the
// bucket holding the k-th key, and the range shrinks >= 2048x; it ends when the width is 1
// (two rounds for costs within a binade or two, three for arbitrary floats).
// A final index-ordered compaction emits every key < T plus the first `take_eq` keys == T:
// thread t owns 16 consecutive indices of each 16384-key pass (selection bit masks, one packed
// block scan), the pass's elite indices are compacted in shared memory and written out coalesced.
// The staged keys are stored with an XOR swizzle of the 16-byte unit index so that both access
// patterns -- thread-strided units (staging, histogram rounds, where the index is irrelevant) and
// four consecutive units per thread (compaction) -- are free of bank conflicts.  Padding keys
// (0xFFFFFFFF up to a multiple of 32) sort after every real key, ties included (highest indices),
// so the rounds and the compaction need no index mask.
//   best (nullable):      (min cost, ., argmin) of this launch per segment
//   best_ever (nullable): updated when this launch's minimum is strictly smaller
//                         (earlier iteration wins ties)
// STAGED=false re-reads the costs from global memory (segments too long for shared memory).
constexpr int kSelectStageMax = 49152;  // keys staged in shared memory: 192 KB
constexpr int kSelectBins = 4096;
constexpr int kSelectBinBits = 12;
__host__ __device__ inline int select_padded(int n) { return (n + 31) & ~31; }

// inclusive warp scan
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, v, d);
    if (lane >= d) v += o;
  }
  return v;
}

template <bool STAGED, int MODE>
__global__ void __launch_bounds__(kSelectThreads)
topk_select_kernel(const float* __restrict__ costs, int n, int k, int* __restrict__ elite_idx,
                   float* __restrict__ elite_cost, MbrlPlanInfo* __restrict__ best,
                   BestEver* __restrict__ best_ever, int iteration, const SelShard sh) {
  extern __shared__ __align__(16) uint32_t sel_smem[];
  uint32_t* keys = sel_smem;  // [select_padded(n)] when STAGED, else the 16384-entry compaction buffer
  __shared__ __align__(16) uint32_t hist[kSelectBins];
  __shared__ uint32_t wtot[2][2][32];  // per-warp totals (double buffered)
  __shared__ uint32_t s_sel[2];        // winning bin, remaining rank
  __shared__ int s_ok;
  __shared__ int s_first;              // lowest index holding the minimum key
  __shared__ uint32_t s_eq_tot, s_eq_low;  // keys == T in all / below this rank's slice (kSelMerge)

  const int seg = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int n32 = select_padded(n);
  // kSelMerge: the gathered candidates live in this rank's exported buffer (written by the peers)
  const size_t par_off = MODE == kSelPlain ? 0 : (size_t)sh.parity * p2p_parity_words(sh.world, sh.slot, sh.pslots);
  const float* c = costs + (long long)seg * n;  // unused by kSelMerge: its costs are packets in sh.local
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(c) & 15) == 0);
  const unsigned long long t_start = MODE == kSelMerge ? globaltimer_ns() : 0ull;
  pdl_trigger();
  // kSelMerge does not wait for the preceding (local select) kernel to drain: everything it consumes
  // arrives as sequence-tagged packets -- its own rank's included, which the local select writes like
  // any other rank's -- and it writes nothing before it holds them.  It is resident and polling while the
  // rollout and the select still run, so the launch gap and the select's kernel end leave the critical path.
  if (MODE != kSelMerge) pdl_wait();  // the costs come from the preceding rollout kernel
  TOPK_STAMP(0);
  SHARD_STAMP(sh.stamps, MODE == kSelMerge ? 2 : 0);
  // kSelMerge: a rank whose packets never show up within timeout_ns (wall clock, default 120 s,
  // MBRL_P2P_TIMEOUT_S) trips the timeout: *error = 1 and every gathered candidate reads as
  // (+inf, -1) instead of stale data, so that whatever runs next is deterministic garbage that the
  // plan reports (info.reserved bit 1), not a plausible wrong plan.
  bool bad = false;
  if (MODE == kSelMerge) {
    if (t == 0) s_ok = 1;
    __syncthreads();
  }
  constexpr uint32_t kInfKey = 0x7F800000u | 0x80000000u;  // cost_key(+inf)

  // four keys of indices i4..i4+3 (i4 multiple of 4); out-of-range -> 0xFFFFFFFF padding
  auto load4 = [&](int i4, uint32_t (&kk)[4]) {
    if (MODE == kSelMerge) {
      const uint32_t* pk = sh.local + par_off + 2 * (size_t)i4;  // cost packets of the gathered indices i4..i4+3
      uint2 q[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) q[j] = (i4 + j < n && !bad) ? ld_packet(pk + 2 * j) : make_uint2(0u, sh.seq);  // in flight together
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        kk[j] = 0xFFFFFFFFu;
        if (i4 + j < n) {
          uint32_t v = q[j].x;
          if (bad || (q[j].y != sh.seq && !wait_packet(pk + 2 * j, sh.seq, t_start, sh.timeout_ns, v))) { s_ok = 0; kk[j] = kInfKey; }
          else kk[j] = cost_key(__uint_as_float(v));
        }
      }
    } else if (vec_ok && i4 + 3 < n) {
      const float4 q = dep_load(reinterpret_cast<const float4*>(c + i4));
      kk[0] = cost_key(q.x); kk[1] = cost_key(q.y); kk[2] = cost_key(q.z); kk[3] = cost_key(q.w);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) kk[j] = i4 + j < n ? cost_key(dep_load(c + i4 + j)) : 0xFFFFFFFFu;
    }
  };
  // staged key of index i lives in unit swz(i / 4): the swizzle permutes units within aligned groups of 8
  auto swz = [](int unit) { return unit ^ ((unit >> 3) & 7); };
  auto key4_at = [&](int i4, uint32_t (&kk)[4]) {  // the four keys of INDICES i4..i4+3
    if (STAGED) {
      const uint4 q = *reinterpret_cast<const uint4*>(keys + 4 * swz(i4 >> 2));
      kk[0] = q.x; kk[1] = q.y; kk[2] = q.z; kk[3] = q.w;
    } else {
      load4(i4, kk);
    }
  };
  auto key4_any = [&](int i4, uint32_t (&kk)[4]) {  // four keys of SOME indices: every unit visited once
    if (STAGED) {
      const uint4 q = *reinterpret_cast<const uint4*>(keys + i4);
      kk[0] = q.x; kk[1] = q.y; kk[2] = q.z; kk[3] = q.w;
    } else {
      load4(i4, kk);
    }
  };

  // ---- stage + min/max (16-byte loads, all in flight together) ----
  uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
  if (MODE == kSelMerge) {
    // Four units per thread (16 packets) are loaded together and the whole batch is re-read until every
    // tag is current: one L2 round trip per attempt, so staging ends within a round trip of the last
    // packet landing.  This kernel is resident long before the data exists (it does not wait for the
    // preceding grids), so until its first packet shows up a thread polls that one packet only.
    for (int b4 = 4 * t; b4 < n32; b4 += 16 * kSelectThreads) {
      uint2 q[4][4];
      bool timed_out = false;
      if (b4 < n) {
        uint32_t v;
        timed_out = !wait_packet(sh.local + par_off + 2 * (size_t)b4, sh.seq, t_start, sh.timeout_ns, v);
      }
      unsigned int spins = 0;
      while (!timed_out) {
        bool all = true;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int i4 = b4 + m * 4 * kSelectThreads;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            q[m][j] = i4 + j < n ? ld_packet(sh.local + par_off + 2 * (size_t)(i4 + j)) : make_uint2(0u, sh.seq);
        }
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
          for (int j = 0; j < 4; ++j) all = all && q[m][j].y == sh.seq;
        if (all) break;
        if ((++spins & 255u) == 0 && globaltimer_ns() - t_start > sh.timeout_ns) timed_out = true;
      }
      if (timed_out) s_ok = 0;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int i4 = b4 + m * 4 * kSelectThreads;
        if (i4 >= n32) break;
        uint32_t kk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          kk[j] = 0xFFFFFFFFu;
          if (i4 + j < n) {
            kk[j] = timed_out ? kInfKey : cost_key(__uint_as_float(q[m][j].x));
            kmin = min(kmin, kk[j]); kmax = max(kmax, kk[j]);
          }
        }
        if (STAGED) *reinterpret_cast<uint4*>(keys + 4 * swz(i4 >> 2)) = make_uint4(kk[0], kk[1], kk[2], kk[3]);
      }
    }
  } else {
#pragma unroll 4
    for (int i4 = 4 * t; i4 < n32; i4 += 4 * kSelectThreads) {
      uint32_t kk[4];
      load4(i4, kk);
      if (STAGED) *reinterpret_cast<uint4*>(keys + 4 * swz(i4 >> 2)) = make_uint4(kk[0], kk[1], kk[2], kk[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (i4 + j < n) { kmin = min(kmin, kk[j]); kmax = max(kmax, kk[j]); }
    }
  }
  *reinterpret_cast<uint4*>(hist + 4 * t) = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xFFFFFFFFu, kmin, d));
    kmax = max(kmax, __shfl_xor_sync(0xFFFFFFFFu, kmax, d));
  }
  if (lane == 0) { wtot[0][0][warp] = kmin; wtot[0][1][warp] = kmax; }
  if (t == 0) { s_first = 0x7FFFFFFF; s_eq_low = 0; }
  __syncthreads();
  if (MODE == kSelMerge) {
    bad = !s_ok;  // block-uniform from here on
    if (bad && t == 0) *sh.error = 1;
    SHARD_STAMP(sh.stamps, 3);
  }
  kmin = wtot[0][0][lane]; kmax = wtot[0][1][lane];  // every warp finishes the reduction itself
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xFFFFFFFFu, kmin, d));
    kmax = max(kmax, __shfl_xor_sync(0xFFFFFFFFu, kmax, d));
  }
  const uint32_t key_min = kmin;  // the argmin is the lowest index holding this key
  uint32_t lo = kmin, hi = kmax, rem = (uint32_t)k;
  TOPK_STAMP(1);

  // k == 1 (random shooting; the last CEM iteration of a plan that keeps no distribution): the threshold is
  // the minimum key itself, no histogram round is needed -- one pass finds the lowest index holding it
  if (k == 1) {
    if (MODE != kSelMerge) {
      for (int i4 = 4 * t; i4 < n32; i4 += 4 * kSelectThreads) {
        uint32_t kk[4];
        key4_any(i4, kk);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (kk[j] == key_min) atomicMin(&s_first, (STAGED ? 4 * swz(i4 >> 2) : i4) + j);
      }
    }
    if (t == 0) s_eq_tot = 0xFFFFFFFFu;  // kSelMerge: always count the ties below this rank's slice
    __syncthreads();
    hi = lo;
  }
  // ---- adaptive radix select: power-of-two bucket width, <= 4096 buckets over [lo, hi] ----
  for (int round = 0; round < (k == 1 ? 0 : 5); ++round) {
    const uint32_t span = hi - lo;                              // in-range test: key - lo <= span
    const int shift = span < (uint32_t)kSelectBins ? 0 : 32 - __clz(span) - kSelectBinBits;  // span >> shift < 4096
    TOPK_STAMP(2 + 3 * round);
    for (int i4 = 4 * t; i4 < n32; i4 += 4 * kSelectThreads) {
      uint32_t kk[4];
      key4_any(i4, kk);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t d = kk[j] - lo;
        if (d <= span) atomicAdd(&hist[d >> shift], 1u);
        // the argmin rides on the first round: d == 0 <=> the minimum key (padding never is: n >= 1)
        if (MODE != kSelMerge && round == 0 && d == 0) atomicMin(&s_first, (STAGED ? 4 * swz(i4 >> 2) : i4) + j);
      }
    }
    __syncthreads();
    TOPK_STAMP(3 + 3 * round);
    // block-wide scan of the 4096 bins: four consecutive bins per thread, warp totals through shared
    // memory, every warp scans the 32 totals itself (no single-warp phase)
    const uint4 m4 = *reinterpret_cast<const uint4*>(hist + 4 * t);
    *reinterpret_cast<uint4*>(hist + 4 * t) = make_uint4(0u, 0u, 0u, 0u);  // ready for the next round
    const uint32_t mine = m4.x + m4.y + m4.z + m4.w;
    const uint32_t incl_w = warp_incl_scan(mine, lane);
    const int buf = round & 1;
    if (lane == 31) wtot[buf][0][warp] = incl_w;
    __syncthreads();
    const uint32_t tot = wtot[buf][0][lane];
    const uint32_t tot_incl = warp_incl_scan(tot, lane);
    const uint32_t base = __shfl_sync(0xFFFFFFFFu, tot_incl - tot, warp);
    const uint32_t incl = base + incl_w, excl = incl - mine;
    if (rem > excl && rem <= incl) {  // exactly one thread: the k-th key is in one of its four bins
      uint32_t r = rem - excl, b = 0;
      if (r > m4.x) { r -= m4.x; b = 1; if (r > m4.y) { r -= m4.y; b = 2; if (r > m4.z) { r -= m4.z; b = 3; } } }
      s_sel[0] = 4u * t + b; s_sel[1] = r;
      s_eq_tot = b == 0 ? m4.x : b == 1 ? m4.y : b == 2 ? m4.z : m4.w;  // == #keys equal to T once the width is 1
    }
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

  if (MODE == kSelScatter && t < sh.world)  // this rank's threshold: the merge proves exactness with it
    st_packet(sh.peers.base[t] + par_off + p2p_rank_off(sh.world, sh.slot) + 6 * (size_t)sh.rank, T, sh.seq);
  if (MODE == kSelMerge && t < sh.world && k_per_rank_lt(n, sh.world, sh.k_full)) {
    // Reduced gather: rank t sent only its k_s = n/world cheapest; its k_s-th key is tl.  tl > T: every
    // candidate it kept back is above the global threshold -> exact.  tl <= T: cheaper-than-threshold
    // candidates of that rank may be missing -> flag, the caller redoes the plan with full-size gathers.
    uint32_t tl = 0;
    if (!bad && !wait_packet(sh.local + par_off + p2p_rank_off(sh.world, sh.slot) + 6 * (size_t)t, sh.seq, t_start, sh.timeout_ns, tl))
      *sh.error = 1;
    if (tl <= T) atomicOr(sh.trunc, 1);
  }

  // ---- index-ordered compaction ----
  // kSelMerge compacts only this rank's slice [own_lo, own_hi) of the gathered candidates: the refit is
  // distributed, every rank needs its own elites alone.  Ties at the threshold are taken in global
  // index order, so the slice must know how many keys == T lie below it -- unless every tied key is an
  // elite anyway (s_eq_tot == take_eq: always, but for exact cost ties straddling the cut).
  const int k_s = MODE == kSelMerge ? n / sh.world : 0;
  const int own_lo = MODE == kSelMerge ? sh.rank * k_s : 0;
  const int own_hi = MODE == kSelMerge ? own_lo + k_s : n32;
  if (MODE == kSelMerge && s_eq_tot != take_eq && own_lo > 0) {  // block-uniform; rare
    uint32_t cnt_low = 0;
    for (int i4 = 4 * t; i4 < n32; i4 += 4 * kSelectThreads) {
      uint32_t kk[4];
      key4_at(i4, kk);
#pragma unroll
      for (int j = 0; j < 4; ++j) cnt_low += (kk[j] == T && i4 + j < own_lo);
    }
    if (cnt_low) atomicAdd(&s_eq_low, cnt_low);
    __syncthreads();
  }
  uint32_t base_less = 0, base_eq = MODE == kSelMerge ? s_eq_low : 0u;
  int pass = 0;
  const int c0 = own_lo & ~31;  // whole swizzle groups: the compaction buffer reuses the consumed key slice
  for (int p0 = c0; p0 < own_hi; p0 += 16 * kSelectThreads, ++pass) {
    const int i0 = p0 + 16 * t;
    uint32_t m_less = 0, m_eq = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t k4[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
      if (i0 + 4 * q < n32) key4_at(i0 + 4 * q, k4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t bit = 1u << (4 * q + j);
        if (k4[j] < T) m_less |= bit;
        if (k4[j] == T) m_eq |= bit;
      }
    }
    if (MODE == kSelMerge) {  // clip to the slice
      uint32_t in = 0xFFFFu;
      if (i0 < own_lo) in &= own_lo - i0 >= 16 ? 0u : 0xFFFFu << (own_lo - i0);
      if (i0 + 16 > own_hi) in &= own_hi - i0 <= 0 ? 0u : 0xFFFFu >> (16 - (own_hi - i0));
      m_less &= in; m_eq &= in;
    }
    TOPK_STAMP(23);
    // one block scan of the packed (less, equal) counts: at most 16384 of either per pass
    const uint32_t nl = __popc(m_less), ne = __popc(m_eq);
    const uint32_t mine = nl | (ne << 16);
    const uint32_t incl_w = warp_incl_scan(mine, lane);
    if (lane == 31) wtot[pass & 1][0][warp] = incl_w;
    __syncthreads();
    const uint32_t tot = wtot[pass & 1][0][lane];
    const uint32_t tot_incl = warp_incl_scan(tot, lane);
    const uint32_t before = __shfl_sync(0xFFFFFFFFu, tot_incl - tot, warp) + incl_w - mine;
    const uint32_t all = __shfl_sync(0xFFFFFFFFu, tot_incl, 31);
    uint32_t eq_before = base_eq + (before >> 16);
    // this pass's elites occupy the output positions [out0, out0 + cnt): positions rise with the index
    const uint32_t eq0 = MODE == kSelMerge ? min(s_eq_low, take_eq) : 0u;  // ties taken below the slice hold no position here
    const uint32_t out0 = base_less + min(base_eq, take_eq) - eq0;
    uint32_t p = base_less + (before & 0xFFFFu) + min(eq_before, take_eq) - eq0 - out0;  // this thread's first position
    base_less += all & 0xFFFFu;
    base_eq += all >> 16;
    const uint32_t cnt = base_less + min(base_eq, take_eq) - eq0 - out0;
    TOPK_STAMP(24);
    // Compact the pass's elite indices in shared memory first (over the pass's own key slice, which
    // every thread has consumed by now), then write them out coalesced: a thread's 16 keys map to
    // scattered positions, and direct stores would cost one 32-byte sector per lane.
    uint32_t* outbuf = STAGED ? keys + p0 : keys;
    if (m_eq == 0) {
      while (m_less) {
        const int j = __ffs(m_less) - 1;
        m_less &= m_less - 1;
        outbuf[p++] = (uint32_t)(i0 + j);
      }
    } else {  // ties at the threshold: the first take_eq of them (by index) belong to the elite set
      uint32_t m = m_less | m_eq;
      while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        const bool e = (m_eq >> j) & 1u;
        if (!e || eq_before < take_eq) outbuf[p++] = (uint32_t)(i0 + j);
        eq_before += e;
      }
    }
    TOPK_STAMP(25);
    __syncthreads();
    TOPK_STAMP(26);
    // four outputs per thread and trip: the gathers (cost / global index of elite i) are in flight together
    for (uint32_t j0 = t; j0 < cnt; j0 += 4 * kSelectThreads) {
      int i[4];
      uint32_t g[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t j = j0 + u * kSelectThreads;
        i[u] = j < cnt ? (int)outbuf[j] : -1;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        g[u] = 0;
        if (i[u] >= 0) {
          if (MODE == kSelScatter) g[u] = __float_as_uint(dep_load(c + i[u]));
          else if (MODE == kSelMerge) {  // this rank's own index packets (written by the local select)
            g[u] = 0xFFFFFFFFu;
            if (!bad && !wait_packet(sh.local + par_off + p2p_idx_off(sh.world, sh.slot) + 2 * (size_t)(i[u] - own_lo), sh.seq, t_start,
                                     sh.timeout_ns, g[u])) { *sh.error = 1; g[u] = 0xFFFFFFFFu; }
          }
          else if (elite_cost) g[u] = __float_as_uint(dep_load(c + i[u]));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (i[u] < 0) continue;
        const uint32_t pos = out0 + j0 + u * kSelectThreads;
        if (MODE == kSelScatter) {
          // peer stores over NVLink: the cost packet into every rank's gathered array, the index packet locally
          const size_t at = par_off + 2 * ((size_t)sh.rank * k + pos);
          for (int r = 0; r < sh.world; ++r) st_packet(sh.peers.base[r] + at, g[u], sh.seq);
          st_packet(sh.peers.base[sh.rank] + par_off + p2p_idx_off(sh.world, sh.slot) + 2 * (size_t)pos,
                    (uint32_t)(i[u] + sh.idx_offset), sh.seq);
        } else if (MODE == kSelMerge) {
          elite_idx[pos] = (int)g[u];
        } else {
          elite_idx[(long long)seg * k + pos] = i[u];
          if (elite_cost) elite_cost[(long long)seg * k + pos] = __uint_as_float(g[u]);
        }
      }
    }
  }
  TOPK_STAMP(21);

  if (MODE == kSelScatter) {
    // this rank's minimum (cost, global index) for the best-ever record, two more packets per peer
    if (t < sh.world) {
      const int v = s_first;
      uint32_t* dst = sh.peers.base[t] + par_off + p2p_rank_off(sh.world, sh.slot) + 6 * (size_t)sh.rank;
      st_packet(dst + 2, __float_as_uint(dep_load(c + v)), sh.seq);
      st_packet(dst + 4, (uint32_t)(v + sh.idx_offset), sh.seq);
    }
    SHARD_STAMP(sh.stamps, 1);
    return;
  }
  if (MODE == kSelMerge) {
    // global minimum = the best of the ranks' minima; equal costs -> the lower rank == the lower global index
    if (t < sh.world) {
      const uint32_t* src_r = sh.local + par_off + p2p_rank_off(sh.world, sh.slot) + 6 * (size_t)t;
      uint32_t cb = 0x7F800000u, gi = 0xFFFFFFFFu;
      if (!bad && !(wait_packet(src_r + 2, sh.seq, t_start, sh.timeout_ns, cb) && wait_packet(src_r + 4, sh.seq, t_start, sh.timeout_ns, gi))) {
        *sh.error = 1; cb = 0x7F800000u; gi = 0xFFFFFFFFu;
      }
      hist[2 * t] = cb; hist[2 * t + 1] = gi;  // the histogram is idle by now
    }
    __syncthreads();
    if (t == 0) {
      uint32_t kb = 0xFFFFFFFFu, cb = 0x7F800000u, gi = 0xFFFFFFFFu;
      for (int r = 0; r < sh.world; ++r) {
        const uint32_t kr = cost_key(__uint_as_float(hist[2 * r]));
        if (r == 0 || kr < kb) { kb = kr; cb = hist[2 * r]; gi = hist[2 * r + 1]; }
      }
      const float cmin = __uint_as_float(cb);
      if (best_ever) {
        BestEver b = best_ever[0];
        if (b.iteration < 0 || cmin < b.cost) { b.cost = cmin; b.iteration = iteration; b.index = (int)gi; best_ever[0] = b; }
      }
      *sh.own_count = (int)(base_less + min(base_eq, take_eq) - min(s_eq_low, take_eq));
      SHARD_STAMP(sh.stamps, 4);
    }
    return;
  }

  // ---- the minimum: the lowest index holding the minimum key (found in the first round) ----
  if (t == 0 && n > 0) {
    const int v = s_first;
    const float cmin = dep_load(c + v);
    if (best) { best[seg].best_cost = cmin; best[seg].best_iteration = iteration; best[seg].best_index = v; best[seg].reserved = 0; }
    if (best_ever) {
      BestEver b = best_ever[seg];
      if (b.iteration < 0 || cmin < b.cost) { b.cost = cmin; b.iteration = iteration; b.index = v; best_ever[seg] = b; }
    }
  }
  SHARD_STAMP(sh.stamps, 1);
  TOPK_STAMP(22);
}

// ---- refit ---------------------------------------------------------------------------
constexpr int kRefitThreads = 1024;
constexpr int kRefitChunk = 2 * kRefitThreads;  // elites per CTA
// CTA width for k elites: one elite per thread up to 1024 (a function of k alone, so the summation
// order -- and with it the bit pattern of the refit -- does not depend on how the population is
// sharded).  Small elite sets (k = 204 at the cfg-5 shard) would otherwise pay for 1024-thread CTAs
// that are 80 % idle: 38 400 of them took 0.8 ms per iteration.
// Several chunks (k > 2048): 512-thread CTAs, four elites per thread -- three to four CTAs fit an SM,
// so that the H*G*chunks CTAs of a large elite set (420 at the 8-way sharded cfg 3) run as one wave.
inline int refit_threads(int k) {
  if (k > kRefitChunk) return kRefitThreads / 2;
  return k >= kRefitThreads ? kRefitThreads : (k < 1 ? 1 : k + 31) / 32 * 32;
}

// Shifted sums over the elites list[e_begin + t], list[e_begin + t + blockDim.x], ... (< e_end) of the
// 4 actions (h, 4g .. 4g+3): each thread regenerates (or gathers) its elites' actions and accumulates
// sum(d), sum(d^2) with d = a - c, c = the old mean of that (h, a) (the draws are centred there, so
// the shifted second moment does not cancel); a fixed shuffle tree + one shared-memory stage reduce
// them.  Threads 0..3 return the sums of action 4g + t in (a1, a2); mu_old is returned for all.
// The order of the additions depends on (the list slice, blockDim.x) only.
__device__ __forceinline__ void refit_accumulate(const ActionSource& src, const Shape& sh, int A, int h, int g, int env_l,
                                                 const int* __restrict__ list, int e_begin, int e_end,
                                                 float (*red)[8], float (&mu_old)[4], float& a1, float& a2) {
  const int G = (A + 3) >> 2;
  const long long R = sh.rows();
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nthreads = blockDim.x;
  const long long ms = ((long long)env_l * sh.H + h) * A;
  const bool inject = src.mode == MBRL_SAMPLE_INJECT_ACTIONS || src.mode == MBRL_SAMPLE_INJECT_NOISE;
  const bool affine = src.mode == MBRL_SAMPLE_INJECT_NOISE || src.mode == MBRL_SAMPLE_GAUSSIAN;
  float sd_old[4], s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ac = min(4 * g + j, A - 1);
    mu_old[j] = affine ? dep_load(src.mu + ms + ac) : 0.f;
    sd_old[j] = affine ? dep_load(src.sd + ms + ac) : 0.f;
  }
  const uint2 key = make_uint2(src.seed_lo, src.seed_hi);
  for (int e = e_begin + t; e < e_end; e += nthreads) {
    const int cand_l = dep_load(list + e);
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
  a1 = 0.f; a2 = 0.f;
  if (t < 4)
    for (int w = 0; w < nthreads / 32; ++w) { a1 += red[w][t]; a2 += red[w][4 + t]; }
}

// mean = c + sum(d)/k, std = sqrt(sum(d^2)/k - (sum(d)/k)^2)   (population std); threads 0..3
__device__ __forceinline__ void refit_finish(int A, int g, long long ms, const float (&mu_old)[4], float a1, float a2, int k,
                                             float* __restrict__ mu_new, float* __restrict__ sd_new) {
  const int t = threadIdx.x;
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

// Parks this CTA's partial sums, counts arrivals; in the last CTA to arrive threads 0..3 return the sum
// of all `nparts` partials IN PART ORDER (-> true), the others return false.
__device__ __forceinline__ bool refit_combine(float* __restrict__ part, unsigned int* __restrict__ arrive, long long slot,
                                              int my_part, int nparts, bool* s_last, float& a1, float& a2) {
  const int t = threadIdx.x;
  float* mine = part + (slot * nparts + my_part) * 8;
  if (t < 4) { mine[t] = a1; mine[4 + t] = a2; }
  __threadfence();
  __syncthreads();
  if (t == 0) {
    const unsigned int n = atomicAdd(arrive + slot, 1u);
    *s_last = n == (unsigned int)nparts - 1;
    if (*s_last) arrive[slot] = 0;  // ready for the next launch (stream order)
  }
  __syncthreads();
  if (!*s_last) return false;
  __threadfence();
  if (t < 4) {
    a1 = 0.f; a2 = 0.f;
    for (int c = 0; c < nparts; ++c) {
      a1 += __ldcg(part + (slot * nparts + c) * 8 + t);
      a2 += __ldcg(part + (slot * nparts + c) * 8 + 4 + t);
    }
  }
  return true;
}

// grid = (H * G, E, chunks): one CTA per (step, 4-wide action group, env, chunk of 2048 elites).
// k > 2048 (large elite sets): the chunk CTAs run in parallel on otherwise idle SMs, park their
// partial sums, and the last one to arrive adds them IN CHUNK ORDER -- the result depends only on
// the elite list (ascending index), never on timing.
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
refit_kernel(ActionSource src, Shape sh, int A, const int* __restrict__ elite_idx, int k,
             float* __restrict__ mu_new, float* __restrict__ sd_new, float* __restrict__ part,
             unsigned int* __restrict__ arrive, long long* stamps) {
  __shared__ float red[kRefitThreads / 32][8];
  __shared__ bool s_last;
  const int G = (A + 3) >> 2;
  const int h = blockIdx.x / G, g = blockIdx.x % G;
  const int env_l = blockIdx.y;
  const int chunk = blockIdx.z, nchunks = gridDim.z;
  const long long ms = ((long long)env_l * sh.H + h) * A;
  float mu_old[4], a1, a2;
  pdl_trigger();
  pdl_wait();  // elite indices (top-k / remap) and the old mean/std (previous refit)
  if (stamps && threadIdx.x == 0 && blockIdx.x + blockIdx.y + blockIdx.z == 0) stamps[2] = (long long)globaltimer_ns();
  refit_accumulate(src, sh, A, h, g, env_l, elite_idx + (long long)env_l * k, chunk * kRefitChunk,
                   min(k, (chunk + 1) * kRefitChunk), red, mu_old, a1, a2);
  if (nchunks > 1 && !refit_combine(part, arrive, (long long)env_l * gridDim.x + blockIdx.x, chunk, nchunks, &s_last, a1, a2)) return;
  refit_finish(A, g, ms, mu_old, a1, a2, k, mu_new, sd_new);
  if (stamps && threadIdx.x == 0 && blockIdx.x + blockIdx.y + blockIdx.z == 0) stamps[3] = (long long)globaltimer_ns();
}

// ---- segment-canonical refit: what a population-sharded run computes ---------------------------
// Population sharding over W ranks distributes the refit: rank r sums ITS OWN elites (candidates
// [r*S, (r+1)*S), S = candidates per rank) with 1024 threads, the W partial sums are exchanged and
// added in rank order.  refit_seg_kernel computes exactly that from the global elite list on one GPU
// (grid.z = W CTAs per slot, segment bounds by binary search in the ascending list): it is the
// refit of the NCCL transport (every rank holds the whole list) and of an UNSHARDED plan asked to
// reproduce a W-way sharded one bit for bit (mbrl_set_refit_segments).
__global__ void __launch_bounds__(kRefitThreads, 1)
refit_seg_kernel(ActionSource src, Shape sh, int A, const int* __restrict__ elite_idx, int k, int seg_size,
                 float* __restrict__ mu_new, float* __restrict__ sd_new, float* __restrict__ part,
                 unsigned int* __restrict__ arrive) {
  __shared__ float red[kRefitThreads / 32][8];
  __shared__ bool s_last;
  __shared__ int s_bound[2];
  const int G = (A + 3) >> 2;
  const int h = blockIdx.x / G, g = blockIdx.x % G;
  const int env_l = blockIdx.y;
  const int s = blockIdx.z, nseg = gridDim.z;
  const long long ms = ((long long)env_l * sh.H + h) * A;
  const int* list = elite_idx + (long long)env_l * k;
  float mu_old[4], a1, a2;
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x < 2) {  // first list position whose candidate index is >= (s + threadIdx.x) * seg_size
    const long long v = (long long)(s + (int)threadIdx.x) * seg_size;
    int lo = 0, hi = k;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if ((long long)dep_load(list + mid) < v) lo = mid + 1; else hi = mid; }
    s_bound[threadIdx.x] = lo;
  }
  __syncthreads();
  refit_accumulate(src, sh, A, h, g, env_l, list, s_bound[0], s_bound[1], red, mu_old, a1, a2);
  if (!refit_combine(part, arrive, (long long)env_l * gridDim.x + blockIdx.x, s, nseg, &s_last, a1, a2)) return;
  refit_finish(A, g, ms, mu_old, a1, a2, k, mu_new, sd_new);
}

// The distributed form over peer memory: grid = H * G CTAs.  Each sums this rank's own elites (list,
// *own_count entries, global candidate indices), stores the 8 partial sums of its slot into every
// rank's buffer as sequence-tagged packets, then polls the same slot's packets of every rank and
// adds the W partials in rank order.  A CTA only ever waits for remote CTAs that publish before
// they wait, so the kernels of different ranks cannot deadlock each other.
struct RefitP2p {
  P2pPeers peers;
  const uint32_t* local;
  int rank, world, slot, pslots, parity;
  uint32_t seq;
  int* error;
  unsigned long long timeout_ns;
  long long* stamps;
};
__global__ void __launch_bounds__(kRefitThreads, 1)
refit_p2p_kernel(ActionSource src, Shape sh, int A, const int* __restrict__ own_list, const int* __restrict__ own_count,
                 int k, float* __restrict__ mu_new, float* __restrict__ sd_new, const RefitP2p px) {
  __shared__ float red[kRefitThreads / 32][8];
  __shared__ float s_pk[64 * 8];  // every rank's 8 partial sums of this slot
  const int G = (A + 3) >> 2;
  const int h = blockIdx.x / G, g = blockIdx.x % G;
  const long long ms = (long long)h * A;
  const int t = threadIdx.x, lane = t & 31;
  float mu_old[4], a1, a2;
  pdl_trigger();
  pdl_wait();
  SHARD_STAMP(px.stamps, 5);
  const int kc = dep_load(own_count);
  refit_accumulate(src, sh, A, h, g, 0, own_list, 0, kc, red, mu_old, a1, a2);
  if (t >= 32) return;
  // Each of the slot's 8 partial sums travels as one naturally aligned 8-byte store {value bits, seq}:
  // the tag arrives with the value, so no system-scope fence and no separate flag are needed.  The
  // parity half that seq selects was last read two iterations ago (see the double-buffering argument
  // in mbrl_b200.h), so a reader can only ever see this iteration's tag or an older one.
  const size_t part = (size_t)px.parity * p2p_parity_words(px.world, px.slot, px.pslots) + p2p_part_off(px.world, px.slot);
  if (t < 8) {
    const float v1 = __shfl_sync(0xFFu, a1, t & 3), v2 = __shfl_sync(0xFFu, a2, t & 3);
    const size_t at = part + ((size_t)px.rank * px.pslots + blockIdx.x) * 16 + 2 * t;  // packet t: sum(d) of action t, or sum(d^2) of action t - 4
    for (int r = 0; r < px.world; ++r) st_packet(px.peers.base[r] + at, __float_as_uint(t < 4 ? v1 : v2), px.seq);
  }
  SHARD_STAMP(px.stamps, 6);
  // the lanes poll the world * 8 packets of this slot (packet q = rank * 8 + j)
  bool ok = true;
  const int npk = px.world * 8;
  const unsigned long long t0 = globaltimer_ns();
  for (int q = lane; q < npk; q += 32) {
    const uint32_t* src_q = px.local + part + ((size_t)(q >> 3) * px.pslots + blockIdx.x) * 16 + 2 * (q & 7);
    uint2 pkt = ld_packet(src_q);
    unsigned int spins = 0;
    while (pkt.y != px.seq) {
      if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > px.timeout_ns) { ok = false; break; }
      pkt = ld_packet(src_q);
    }
    s_pk[q] = __uint_as_float(pkt.x);
  }
  ok = __all_sync(0xFFFFFFFFu, ok);
  __syncwarp();
  SHARD_STAMP(px.stamps, 7);
  if (!ok && t == 0) *px.error = 1;  // the plan reports it (info.reserved bit 1); the sums below are garbage then
  if (t < 4) {
    a1 = 0.f; a2 = 0.f;
    for (int r = 0; r < px.world; ++r) { a1 += s_pk[r * 8 + t]; a2 += s_pk[r * 8 + 4 + t]; }
  }
  refit_finish(A, g, ms, mu_old, a1, a2, k, mu_new, sd_new);
  SHARD_STAMP(px.stamps, 8);
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

// The initial states of a host-buffer plan: read from the handle's pinned, device-mapped host buffer by
// the first kernel of the plan instead of a host-to-device copy in front of it (one PCIe read round).
struct StageS0 {
  const float* src;  // null: nothing to stage
  float* dst;
  int n;
};
__device__ __forceinline__ void stage_s0(const StageS0& s) {
  if (!s.src) return;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < s.n; i += stride)
    s.dst[i] = *reinterpret_cast<const volatile float*>(s.src + i);
}

// mu/sd initialisation: (lo+hi)/2 and (hi-lo)/2 per (env, h, a); best-ever reset.
__global__ void init_plan_kernel(float* __restrict__ mu, float* __restrict__ sd, long long n,
                                 float lo, float hi, BestEver* __restrict__ best_ever, int E, const StageS0 s0) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  pdl_trigger();
  pdl_wait();
  stage_s0(s0);
  if (mu && i < n) { mu[i] = 0.5f * (lo + hi); sd[i] = 0.5f * (hi - lo); }
  if (best_ever && i < E) best_ever[i] = BestEver{0.f, -1, -1, 0};
}

// Warm start resident on the device (MBRL_WARM_USE): the new mean is the previous plan's final mean
// shifted by one step with the last step repeated (the MPC receding-horizon shift of
// MPCPolicy's hand-over, src/mbrl/agents.py:41-47), std = `std`; best-ever reset.
__global__ void warm_init_kernel(float* __restrict__ mu, float* __restrict__ sd, const float* __restrict__ last_mu,
                                 int E, int H, int A, float std, float lo, float hi, BestEver* __restrict__ best_ever,
                                 const StageS0 s0) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  pdl_trigger();
  pdl_wait();
  stage_s0(s0);
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
