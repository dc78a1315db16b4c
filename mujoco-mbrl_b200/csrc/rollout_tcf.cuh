// Fused-recurrence tcgen05 rollout engine (default for MBRL_ENGINE_TC_*; rollout_tc.cuh is
// the unfused fallback for shapes whose fused operands do not fit shared memory).
//
// The reference recurrence per step (src/mbrl/models.py:13-29, 106-110) is
//     x = [norm(s), norm(a)],  h1 = relu(W1 x + b1),  h2 = relu(W2 h1 + b2),  y = W3 h2 + b3,
//     s' = unnorm(y),  next x = [y, norm(a')]        (normalise o unnormalise == identity)
// There is no nonlinearity between layer 3 of step h and layer 1 of step h+1, so
//     h1' = relu(W1s y + W1a a' + b1) = relu((W1s W3) h2 + W1a a' + (b1 + W1s b3)).
// The host packs W13 = W1s*W3 (U x U) and b13 = b1 + W1s*b3 once, and every step becomes TWO
// chained GEMMs on the tensor cores instead of three plus a serial small-GEMM/epilogue chain:
//     GEMM-A(h+1): D_A[128 x (Np+Oy)] = a'(h+1) . [W1a|b13|0]^T  (SS, 1 K-step, actions + const 1)
//                                     + h2(h)   . [W13 ; W3]^T   (TS, A operand from TMEM)
//                  columns [0,Np)      -> pre-activation of h1(h+1)
//                  columns [Np,Np+Oy)  -> y(h) = W3 h2(h): consumed OFF the critical path by the
//                                         cost threads (fp32: + b3, un-normalise, SmoothAbs cost)
//     GEMM-B(h):   D_B[128 x Np] = h1(h) . W2p^T                  (TS)
// Both hidden epilogues (TMEM fp32 -> relu -> 16 bit -> TMEM, each 16-column K-step packed into the
// first half of its own columns, so no warp ever overwrites data another warp still has to read)
// release the next GEMM's K-steps in groups of 4 through mbarriers, so the tensor pipe only idles for the first-group
// latency of each epilogue.  Step 0 uses a state tile (x0 - b3, so that b13 gives exactly b1)
// with W1s as extra SS K-steps.
//
// Warp roles (800 threads, 1 CTA/SM): warps 0-15 hidden epilogues (four warpgroups; warpgroup g
// owns the 16-column K-steps ks = g mod 4, each warp its TMEM lane quarter); then 4 warps cost
// epilogue, one thread per candidate row (fp32: y + b3, un-normalise, SmoothAbs); 4 warps action
// sampler, one thread per row (Philox + Box-Muller + clip, up to 3 steps ahead, Cosh cost); the
// last warp: weight TMA + all tcgen05.mma issue (one elected lane).
// The per-element epilogue math is written branch-free (zero-padded tables and masks): per-element
// branches serialise the load -> fma -> sqrt chains and cost ~10x (measured).
#pragma once
#include <algorithm>
#include <cstdlib>

#include "rollout_tc.cuh"
#include "rollout_tcw.cuh"

namespace mbrl {

struct TcfGeom {
  int O, A, U;
  int Ka;  // action K-range (multiple of 16): A actions, constant 1, zero pad
  int Ks;  // state K-range of the step-0 tile (multiple of 16)
  int Np;  // padded hidden width (multiple of 16, > U)
  int Oy;  // y columns (multiple of 16)
  int Na;  // Np + Oy: N of GEMM-A
  int wa_off, waa_off, w1s_off, w2_off, w_bytes;
  int tab_off, xs_off, xa_off, bar_off, smem_bytes;
  int ms_off, ms_floats;  // staging area for the sampling mean/std rows of the tile's environments
  int exp;      // profiling experiments (MBRL_TCF_EXP bit mask)
  int xch_off;  // task costs: per-step control terms (a0, ctl_mean) handed from the sampler to the cost threads; -1 = no room
  long long* stamps;  // diagnostic (MBRL_PLAN_TIMELINE): per CTA globaltimer at entry / upstream data ready / exit, or null
};

constexpr int kTcfSlots = 3;  // action tiles in flight: the sampler runs up to 3 steps ahead
constexpr int kTcfXchSlots = 8;  // control-term slots (the cost threads lag the sampler by at most 4 steps)
constexpr int kTcfMaxKSteps = 16;  // hidden K-steps (Np <= 256)
constexpr int kTcfBarriers = 4 + 2 * kTcfSlots + kTcfMaxKSteps;  // w, dA, dB, y, xa[3], xf[3], hA[8], hB[8]
constexpr int kTcfEpiGroups = 4;                      // hidden-epilogue warpgroups (4 warps each)
constexpr int kTcfCostWarp0 = 4 * kTcfEpiGroups;      // 4 warps: cost epilogue (TMEM lane quarter = warp & 3)
constexpr int kTcfSampWarp0 = kTcfCostWarp0 + 4;      // 4 warps: action sampler
constexpr int kTcfMmaWarp = kTcfSampWarp0 + 4;        // 1 warp: TMEM alloc, weight TMA, MMA issue
constexpr int kTcfThreads = (kTcfMmaWarp + 1) * 32;

inline bool tcf_geometry(int O, int A, int U, size_t max_smem, TcfGeom* g, std::string* why) {
  g->O = O; g->A = A; g->U = U;
  g->Ka = round_up(A + 1, 16);
  g->Ks = round_up(O, 16);
  g->Np = round_up(U + 1, 16);
  g->Oy = round_up(O, 16);
  g->Na = g->Np + g->Oy;
  if (g->Na > 256) { *why = "hidden + obs too wide for the fused GEMM (N > 256)"; return false; }
  g->wa_off = 0;
  g->waa_off = g->wa_off + g->Np * g->Na * 2;
  g->w1s_off = g->waa_off + g->Ka * g->Na * 2;
  g->w2_off = g->w1s_off + g->Ks * g->Np * 2;
  g->w_bytes = g->w2_off + g->Np * g->Np * 2;
  g->tab_off = g->w_bytes;  // fp32: 6 tables of Oy, 2 of kMaxAct, 2 x 128 cost partials
  g->xs_off = round_up(g->tab_off + (6 * g->Oy + 2 * kMaxAct + 2 * kTcRows) * 4, 128);
  g->xa_off = g->xs_off + g->Ks * kTcRows * 2;
  g->bar_off = g->xa_off + kTcfSlots * g->Ka * kTcRows * 2;
  g->smem_bytes = g->bar_off + 8 * kTcfBarriers + 16;
  if ((size_t)g->smem_bytes > max_smem) { *why = "fused operands do not fit shared memory"; return false; }
  // dm_control task costs that depend on the control (cartpole, humanoid): the sampler hands (a0,
  // ctl_mean) of every step to the cost threads through kTcfXchSlots slots -- only when shared memory
  // has the room (small models; the cheetah / walker shapes do not need it: their task costs read the
  // state only)
  g->xch_off = -1;
  g->exp = 0;
  const int xch_bytes = kTcfXchSlots * 2 * kTcRows * 4;
  if ((size_t)g->smem_bytes + xch_bytes + 4096 <= max_smem) { g->xch_off = g->smem_bytes; g->smem_bytes += xch_bytes; }
  // whatever is left (up to 16 KB) stages mean/std: [envs of the tile][H][A] x 2, fp32
  g->ms_off = g->smem_bytes;
  g->ms_floats = (int)std::min<size_t>((max_smem - (size_t)g->smem_bytes) / 4, 4096);
  g->smem_bytes += 4 * g->ms_floats;
  return true;
}

// Packs [W13;W3], [W1a|b13], W1s, W2p as 16-bit canonical K-major operands (see tc_put).
inline void tcf_pack(const TcfGeom& g, bool fp16, const float* W1, const float* b1, const float* W2, const float* b2,
                     const float* W3, const float* b3, std::vector<uint16_t>* out) {
  const int O = g.O, A = g.A, U = g.U, D = O + A;
  std::vector<uint16_t>& img = *out;
  img.assign((size_t)g.w_bytes / 2, 0);
  std::vector<double> w13((size_t)U * U), b13(U);
  for (int n = 0; n < U; ++n) {
    double b = b1[n];
    for (int o = 0; o < O; ++o) b += (double)W1[(size_t)n * D + o] * b3[o];
    b13[n] = b;
    for (int k = 0; k < U; ++k) {
      double acc = 0.0;
      for (int o = 0; o < O; ++o) acc += (double)W1[(size_t)n * D + o] * W3[(size_t)o * U + k];
      w13[(size_t)n * U + k] = acc;
    }
  }
  for (int n = 0; n < U; ++n)
    for (int k = 0; k < U; ++k) tc_put(img, g.wa_off, g.Na, n, k, (float)w13[(size_t)n * U + k], fp16);
  for (int o = 0; o < O; ++o)
    for (int k = 0; k < U; ++k) tc_put(img, g.wa_off, g.Na, g.Np + o, k, W3[(size_t)o * U + k], fp16);
  for (int n = 0; n < U; ++n) {
    for (int a = 0; a < A; ++a) tc_put(img, g.waa_off, g.Na, n, a, W1[(size_t)n * D + O + a], fp16);
    tc_put(img, g.waa_off, g.Na, n, A, (float)b13[n], fp16);
  }
  tc_put(img, g.waa_off, g.Na, U, A, 1.0f, fp16);  // hidden unit U == relu(1) == 1 carries b2
  for (int n = 0; n < U; ++n)
    for (int o = 0; o < O; ++o) tc_put(img, g.w1s_off, g.Np, n, o, W1[(size_t)n * D + o], fp16);
  for (int n = 0; n < U; ++n) {
    for (int k = 0; k < U; ++k) tc_put(img, g.w2_off, g.Np, n, k, W2[(size_t)n * U + k], fp16);
    tc_put(img, g.w2_off, g.Np, n, U, b2[n], fp16);
  }
}

// DBG: accumulator dump + clock64 timeline instrumentation (tests / profiling only); the
// production instantiation carries none of it.
// SPEC: a geometry class as compile-time constants, so that the single MMA-issuing thread runs
// straight-line code with immediate operand offsets (see rollout_tcw.cuh for the measurements) and the
// sampler's per-action clamps and masks fold away.  0 = run-time geometry;
//   1 = BASELINE cfgs 3 and 4 (cheetah-run O=17 / walker-walk O=24, A=6, hidden 200): Np 208, Oy 32, Ka 16, Ks 32
//   2 = BASELINE cfgs 1 and 2 (cartpole-swingup O=5, A=1, hidden 50):                 Np  64, Oy 16, Ka 16, Ks 16
constexpr int kTcfSpecs = 2;
__host__ __device__ constexpr int tcf_spec_np(int s) { return s == 1 ? 208 : 64; }
__host__ __device__ constexpr int tcf_spec_oy(int s) { return s == 1 ? 32 : 16; }
__host__ __device__ constexpr int tcf_spec_ka(int s) { return 16; }
__host__ __device__ constexpr int tcf_spec_ks(int s) { return s == 1 ? 32 : 16; }
__host__ __device__ constexpr int tcf_spec_a(int s) { return s == 1 ? 6 : 1; }
inline int tcf_matches_spec(const TcfGeom& g) {  // the matching class, 0 = none
  for (int s = 1; s <= kTcfSpecs; ++s)
    if (g.Np == tcf_spec_np(s) && g.Oy == tcf_spec_oy(s) && g.Ka == tcf_spec_ka(s) && g.Ks == tcf_spec_ks(s) && g.A == tcf_spec_a(s))
      return s;
  return 0;
}

// TASK: a dm_control task cost instead of SmoothAbs + Cosh (compiled out of the default-cost kernel: the
// extra registers and branches in the sampler / cost threads cost the default path 12 % when left in).
template <bool FP16, bool DBG, int SPEC, bool TASK>
__global__ void __launch_bounds__(kTcfThreads, 1)
rollout_tcf_kernel(TcfGeom g, const uint8_t* __restrict__ wimg, ModelDev m, ActionSource src, Shape sh,
                   const float* __restrict__ s0, float* __restrict__ costs, float* __restrict__ states_out,
                   float* __restrict__ actions_out, float* __restrict__ dbg) {
  extern __shared__ __align__(128) uint8_t tcf_smem[];
  uint8_t* const smem = tcf_smem;
  // PDL: barrier init, TMEM allocation, the 207 KB weight TMA and the table fill only touch static
  // model data, so they overlap the tail of the preceding refit / top-k.  The sampler and cost
  // threads are the only consumers of upstream results (mean/std, s0) and wait below; the final
  // cost write follows a block barrier that those threads have passed.
  pdl_trigger();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (g.stamps && tid == 0) g.stamps[3 * blockIdx.x] = (long long)globaltimer_ns();
  const int O = g.O, A = g.A, H = sh.H;
  const int NpC = SPEC ? tcf_spec_np(SPEC) : g.Np, OyC = SPEC ? tcf_spec_oy(SPEC) : g.Oy, NaC = NpC + OyC;
  constexpr bool smooth = !TASK;
  float* const xch = (TASK && g.xch_off >= 0) ? reinterpret_cast<float*>(smem + g.xch_off) : nullptr;  // [slot][a0, ctl][row]
  const int KS_H = NpC >> 4;        // K-steps over a hidden operand
  const int KS_A = SPEC ? tcf_spec_ka(SPEC) >> 4 : g.Ka >> 4, KS_S = SPEC ? tcf_spec_ks(SPEC) >> 4 : g.Ks >> 4;

  float* tab = reinterpret_cast<float*>(smem + g.tab_off);
  float *t_b3 = tab, *t_P = tab + g.Oy, *t_Q = tab + 2 * g.Oy, *t_sd = tab + 3 * g.Oy, *t_mu = tab + 4 * g.Oy;
  float *t_M = tab + 5 * g.Oy;  // 1 for real outputs, 0 for padding
  float *t_ainv = tab + 6 * g.Oy, *t_aoff = t_ainv + kMaxAct, *costp = t_aoff + kMaxAct;
  uint8_t* xs = smem + g.xs_off;
  uint8_t* xa = smem + g.xa_off;
  const int xa_bytes = g.Ka * kTcRows * 2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.bar_off);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + g.bar_off + 8 * kTcfBarriers);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t bar_w = bar0, bar_dA = bar0 + 8, bar_dB = bar0 + 16, bar_y = bar0 + 24, bar_xa = bar0 + 32;
  const uint32_t bar_xf = bar_xa + 8 * kTcfSlots;  // action tile slot free again (its MMAs completed)
  const uint32_t bar_hA = bar_xf + 8 * kTcfSlots, bar_hB = bar_hA + 2 * kTcfMaxKSteps;

  if (warp == kTcfMmaWarp) {
    if (lane == 0) {
      mbar_init(bar_w, 1); mbar_init(bar_dA, 1); mbar_init(bar_dB, 1); mbar_init(bar_y, 1);
      for (int i = 0; i < kTcfSlots; ++i) { mbar_init(bar_xa + 8 * i, 1); mbar_init(bar_xf + 8 * i, 1); }
      for (int c = 0; c < kTcfMaxKSteps / 4; ++c) {
        // one barrier per group of 4 K-steps (the 4 warpgroups convert one K-step each, in
        // parallel): 4 warp arrivals per K-step.  Coarser than a pair so that the issue loop's
        // fixed cost (~290 cycles per barrier wait + fence + elect) is paid once per 4 MMAs.
        const int ksteps = min(4, max(0, (g.Np >> 4) - 4 * c));
        if (ksteps > 0) { mbar_init(bar_hA + 8 * c, 4 * ksteps); mbar_init(bar_hB + 8 * c, 4 * ksteps); }
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < g.Oy; i += kTcfThreads) {
    const bool in = i < O;
    t_M[i] = in ? 1.f : 0.f;
    const float b3 = in ? __ldg(m.b3 + i) : 0.f, sd = in ? __ldg(m.sd_s + i) : 1.f, mu = in ? __ldg(m.mu_s + i) : 0.f;
    const float w = (in && smooth) ? __ldg(m.cost_w + i) : 0.f, goal = (in && smooth) ? __ldg(m.goal + i) : 0.f;
    t_b3[i] = b3; t_sd[i] = sd; t_mu[i] = mu;
    t_P[i] = sd * w;
    t_Q[i] = (b3 * sd + mu - goal) * w;
  }
  for (int i = tid; i < kMaxAct; i += kTcfThreads) {
    // normalised action = a * inv - off; the constant-1 column (i == A) is 0 * 0 - (-1)
    const float inv = i < A ? 1.0f / __ldg(m.sd_a + i) : 0.f;
    t_ainv[i] = inv;
    t_aoff[i] = i < A ? __ldg(m.mu_a + i) * inv : (i == A ? -1.f : 0.f);
  }
  // zero the action tiles once: chunks beyond the sampled ones stay zero for the whole rollout
  for (int i = tid; i < kTcfSlots * xa_bytes / 16; i += kTcfThreads) reinterpret_cast<uint4*>(xa)[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const long long R = sh.rows();

  if (warp == kTcfMmaWarp) {
    // ================= MMA issuer: ONE elected lane runs the whole loop =================
    // tcgen05.mma issue blocks the issuing thread for about the MMA's own duration, so the thread's own
    // bookkeeping between MMAs is tensor-pipe idle time.  The loops below are written for full unrolling
    // in the SPEC instantiation (compile-time geometry of the cheetah / walker shape class): descriptors
    // as (lo, hi) words whose start-address field takes immediate offsets, the TMEM base as the literal
    // 0 (a 512-column allocation starts there; checked), no per-group elect / __syncwarp.
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_w, (uint32_t)g.w_bytes);
      bulk_g2s(smem_u32(smem), wimg, (uint32_t)g.w_bytes, bar_w);
      if (tmem != 0u) __trap();
      const uint32_t idesc_a = umma_idesc(NaC, FP16), idesc_h = umma_idesc(NpC, FP16), idesc_y = umma_idesc(OyC, FP16);
      const uint32_t dhi = (128u >> 4) | (1u << 14);  // SBO = 128 bytes; descriptor version 1 (bit 46)
      auto dlo = [](uint32_t saddr, uint32_t lbo_bytes) { return ((saddr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16); };
      auto d64 = [dhi](uint32_t lo) { return ((uint64_t)dhi << 32) | lo; };
      const uint32_t lbo_a = (uint32_t)NaC * 16, lbo_h = (uint32_t)NpC * 16, lbo_x = kTcRows * 16;
      // a K-step advances the 14-bit start-address field by 2*LBO/16 (smem addresses stay below 256 KB)
      const uint32_t lo_wa = dlo(smem_u32(smem + g.wa_off), lbo_a);
      const uint32_t lo_wy = dlo(smem_u32(smem + g.wa_off) + (uint32_t)NpC * 16, lbo_a);  // rows Np.. of [W13;W3]
      const uint32_t lo_waa = dlo(smem_u32(smem + g.waa_off), lbo_a);
      const uint32_t lo_w1s = dlo(smem_u32(smem + g.w1s_off), lbo_h);
      const uint32_t lo_w2 = dlo(smem_u32(smem + g.w2_off), lbo_h);
      const uint32_t lo_xs = dlo(smem_u32(xs), lbo_x), lo_xa0 = dlo(smem_u32(xa), lbo_x);
      const uint32_t xa_step = (uint32_t)(xa_bytes >> 4);  // next action tile
      const uint32_t step_a = (2 * lbo_a) >> 4, step_h = (2 * lbo_h) >> 4, step_x = (2 * lbo_x) >> 4;
      const uint32_t tm_a = 0u, tm_b = kTcD2Col;
      mbar_wait(bar_w, 0);

      // ---- step 0: D_A = a(0).[W1a|b13]^T + (x0 - b3).W1s^T ----
      mbar_wait(bar_xa, 0);
      tc_fence_after();
#pragma unroll (SPEC ? 4 : 1)
      for (int ks = 0; ks < KS_A; ++ks) mma_ss(tm_a, d64(lo_xa0 + ks * step_x), d64(lo_waa + ks * step_a), idesc_a, ks > 0);
      tc_commit(bar_xf);  // slot 0 is free again once these MMAs have read it
#pragma unroll (SPEC ? 4 : 1)
      for (int ks = 0; ks < KS_S; ++ks) mma_ss(tm_a, d64(lo_xs + ks * step_x), d64(lo_w1s + ks * step_h), idesc_h, 1);
      tc_commit(bar_dA);

      uint32_t slot = 0, slot_ph = 0;  // action-tile slot of step h+1 and its phase
      for (int h = 0; h < H; ++h) {
        const uint32_t ph = h & 1;
        // ---- GEMM-B(h): D_B = h1(h) . W2p^T, K-steps released by epilogue A ----
        // K-step ks of h1 (16 hidden units) is packed into the first 8 of its own 16 fp32 columns; a
        // group of 4 K-steps is released on its barrier as soon as the four warpgroups have converted it
#pragma unroll (SPEC ? 4 : 1)
        for (int ks = 0; ks < KS_H; ks += 4) {
          if (g.exp & 2) mbar_wait_nohint(bar_hA + 2 * ks, ph); else mbar_wait(bar_hA + 2 * ks, ph);  // barrier of the K-step group ks/4
          tc_fence_after();
          if (DBG && (ks == 0 || ks + 4 >= KS_H)) tc_stamp(dbg, h, ks == 0 ? 1 : 2);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (ks + i < KS_H) mma_ts(tm_b, tm_a + 16 * (ks + i), d64(lo_w2 + (ks + i) * step_h), idesc_h, (ks + i) > 0);
        }
        if (DBG) tc_stamp(dbg, h, 8);
        tc_commit(bar_dB);
        if (DBG) tc_stamp(dbg, h, 9);
        // ---- GEMM-A(h+1) (or, after the last step, only the y columns) ----
        const bool last = h + 1 == H;
        if (h >= 1) mbar_wait(bar_y, (h - 1) & 1);  // y(h-1) in D_A has been consumed
        if (DBG) tc_stamp(dbg, h, 10);
        uint32_t acc = 0;
        if (!last) {
          if (++slot == kTcfSlots) { slot = 0; slot_ph ^= 1; }  // slot (h+1) % kTcfSlots, phase ((h+1) / kTcfSlots) & 1
          mbar_wait(bar_xa + 8 * slot, slot_ph);
          tc_fence_after();
          if (DBG) tc_stamp(dbg, h, 3);
          const uint32_t lo_xa = lo_xa0 + slot * xa_step;
#pragma unroll (SPEC ? 4 : 1)
          for (int ks = 0; ks < KS_A; ++ks) mma_ss(tm_a, d64(lo_xa + ks * step_x), d64(lo_waa + ks * step_a), idesc_a, ks > 0);
          tc_commit(bar_xf + 8 * slot);
          acc = 1;
        }
        {
          const uint32_t lo_b = last ? lo_wy : lo_wa;
          const uint32_t d = last ? tm_a + (uint32_t)NpC : tm_a;
          const uint32_t idesc = last ? idesc_y : idesc_a;
          if (DBG) tc_stamp(dbg, h, 18);
#pragma unroll (SPEC ? 4 : 1)
          for (int ks = 0; ks < KS_H; ks += 4) {
            if (g.exp & 2) mbar_wait_nohint(bar_hB + 2 * ks, ph); else mbar_wait(bar_hB + 2 * ks, ph);
            tc_fence_after();
            if (DBG && (ks == 0 || ks + 4 >= KS_H)) tc_stamp(dbg, h, ks == 0 ? 16 : 17);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (ks + i < KS_H) mma_ts(d, tm_b + 16 * (ks + i), d64(lo_b + (ks + i) * step_a), idesc, (ks + i) > 0 ? 1u : acc);
          }
          tc_commit(bar_dA);
        }
        if (DBG) tc_stamp(dbg, h + 1, 0);
      }
    }
    __syncwarp();
  } else if (warp >= kTcfSampWarp0) {
    // ================= sampler threads (one per row) =================
    const int srow = tid - kTcfSampWarp0 * 32;
    const long long row = (long long)blockIdx.x * kTcRows + srow;
    const bool valid = row < R;
    const int env_l = valid ? (int)(row / sh.N) : 0;
    const int cand_l = valid ? (int)(row - (long long)env_l * sh.N) : 0;
    // The sampler warps are the kernel's critical path (one row per thread, everything serial): with the
    // action count a compile-time constant the index clamps, masks and padding lanes fold away.
    const int AC = SPEC ? tcf_spec_a(SPEC) : A;
    const float inv_beta = 1.0f / m.beta, cscale = (valid && smooth) ? m.beta2 / (float)AC : 0.f;
    const int QA = (AC + 8) >> 3;  // 8-wide chunks holding the actions and the constant 1
    float act_total = 0.f;
    // (Drawing step 0's noise before this wait was tried: -2 % -- the extra live registers spill.)
    pdl_wait();
    if (g.stamps && srow == 0) g.stamps[3 * blockIdx.x + 1] = (long long)globaltimer_ns();

    // The mean/std of the sampling distribution (written by the preceding refit) are re-read every
    // step by every row: copy the rows of this tile's environments into shared memory once
    // (coherent loads), so the per-step reads are not L2 round trips on the sampler's chain.
    const bool gauss = src.mode == MBRL_SAMPLE_INJECT_NOISE || src.mode == MBRL_SAMPLE_GAUSSIAN;
    const long long row_lo = (long long)blockIdx.x * kTcRows, row_hi = min(row_lo + kTcRows, R) - 1;
    const int env_lo = (int)(row_lo / sh.N), env_hi = (int)(row_hi / sh.N);
    const int ms_n = (env_hi - env_lo + 1) * H * A;
    const bool staged = gauss && 2 * ms_n <= g.ms_floats;  // uniform over the CTA
    float* const ms_mu = reinterpret_cast<float*>(smem + g.ms_off);
    float* const ms_sd = ms_mu + ms_n;
    if (staged) {
      const long long base = (long long)env_lo * H * A;
      for (int i = srow; i < ms_n; i += kTcRows) { ms_mu[i] = dep_load(src.mu + base + i); ms_sd[i] = dep_load(src.sd + base + i); }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    const int ms_env = (env_l - env_lo) * H * AC;

    auto stage_actions = [&](int hs) {
      float acc = 0.f;
      float* aout = (actions_out && valid) ? actions_out + ((long long)hs * R + row) * AC : nullptr;
      uint8_t* xt = xa + (hs % kTcfSlots) * xa_bytes;
      for (int q = 0; q < QA; ++q) {
        float v[8];
        {
          float t4[4], u4[4];
          const float* pm = staged ? ms_mu + ms_env + hs * AC : nullptr;
          const float* ps = staged ? ms_sd + ms_env + hs * AC : nullptr;
          raw_action4(src, AC, H, hs, env_l, cand_l, row, R, 2 * q, t4, pm, ps);
          if (8 * q + 4 < AC) raw_action4(src, AC, H, hs, env_l, cand_l, row, R, 2 * q + 1, u4, pm, ps);
          else { u4[0] = u4[1] = u4[2] = u4[3] = 0.f; }
          v[0] = t4[0]; v[1] = t4[1]; v[2] = t4[2]; v[3] = t4[3];
          v[4] = u4[0]; v[5] = u4[1]; v[6] = u4[2]; v[7] = u4[3];
        }
        // raw_action4 returns 0 beyond A (cosh(0) - 1 == 0), tables are zero-padded: lanes beyond a
        // compile-time A are dropped, a run-time A runs all 8 branch-free
        if (xch && !smooth) {
          // task costs: this step's first control and mean_a(quadratic tolerance of the control), accumulated
          // in the step's exchange slot (no live registers on the default-cost path)
          float c8 = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) c8 += (8 * q + i < AC && fabsf(v[i]) < 1.0f) ? 1.0f - v[i] * v[i] : 0.0f;
          float* slot = xch + ((hs % kTcfXchSlots) * 2) * kTcRows + srow;
          if (q == 0) { slot[0] = v[0]; slot[kTcRows] = c8 / (float)AC; }
          else slot[kTcRows] += c8 / (float)AC;
        }
        float xn[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (!SPEC || 8 * q + i < AC) acc += cosh_m1_fast(v[i] * inv_beta);
          xn[i] = fmaf(v[i], t_ainv[8 * q + i], -t_aoff[8 * q + i]);  // the lane after the last action holds the constant 1
        }
        if (aout) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (8 * q + i < AC) aout[8 * q + i] = v[i];
        }
        uint4 pk;
        pk.x = pack2<FP16>(xn[0], xn[1]); pk.y = pack2<FP16>(xn[2], xn[3]);
        pk.z = pack2<FP16>(xn[4], xn[5]); pk.w = pack2<FP16>(xn[6], xn[7]);
        *reinterpret_cast<uint4*>(xt + q * (kTcRows * 16) + srow * 16) = pk;
      }
      act_total = fmaf(cscale, acc, act_total);  // CoshLoss: beta^2 * mean_a(cosh(a/beta) - 1)
    };

    // step-0 tiles: actions(0) and the state tile x0 - b3 (so that b13 reduces to b1)
    for (int j = 0; j < (g.Ks >> 3); ++j) {
      float xn[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int o = 8 * j + i;
        xn[i] = (o < O && valid) ? (dep_load(s0 + (long long)env_l * O + o) - t_mu[o]) / t_sd[o] - t_b3[o] : 0.f;
      }
      uint4 pk;
      pk.x = pack2<FP16>(xn[0], xn[1]); pk.y = pack2<FP16>(xn[2], xn[3]);
      pk.z = pack2<FP16>(xn[4], xn[5]); pk.w = pack2<FP16>(xn[6], xn[7]);
      *reinterpret_cast<uint4*>(xs + j * (kTcRows * 16) + srow * 16) = pk;
    }
    for (int hs = 0; hs < H; ++hs) {
      // Tile slot hs%3 was last read by the action MMAs of step hs-3; their commit on the slot's
      // own "free" barrier is phase #(hs/3 - 1) -- the waiter is never more than one phase behind
      // (a parity wait cannot tell phases that are 2 apart).
      if (hs >= kTcfSlots) mbar_wait(bar_xf + 8 * (hs % kTcfSlots), (hs / kTcfSlots - 1) & 1);
      if (srow == 0) if (DBG) tc_stamp(dbg, hs, 12);
      stage_actions(hs);
      fence_proxy_async();   // generic-proxy tile writes -> visible to the MMA (async proxy)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (srow == 0) { mbar_arrive(bar_xa + 8 * (hs % kTcfSlots)); if (DBG) tc_stamp(dbg, hs, 13); }
    }
    costp[kTcRows + srow] = act_total;
  } else if (warp >= kTcfCostWarp0) {
    // ================= cost threads (one per row) =================
    const int crow = tid - kTcfCostWarp0 * 32;
    const long long row = (long long)blockIdx.x * kTcRows + crow;
    const bool valid = row < R;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float st_total = 0.f;
    pdl_wait();
    // Phases must be observed in order: a parity wait on phase #1 issued before phase #0 has
    // completed would fall through at once (it cannot tell "not yet" from "one phase ago").
    mbar_wait(bar_dA, 0);
    for (int j = 1; j <= H; ++j) {
      // D_A(j) carries y(j-1) = W3 h2(j-1) in columns [Np, Np+Oy)  (j == H: the y-only GEMM)
      mbar_wait(bar_dA, j & 1);
      tc_fence_after();
      if (crow == 0) if (DBG) tc_stamp(dbg, j - 1, 14);
      float* sout = (states_out && valid) ? states_out + ((long long)(j - 1) * R + row) * O : nullptr;
      float p4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int cc = 0; cc < (OyC >> 4); ++cc) {
        uint32_t v[32];
        tmem_ld16(lane_base + (uint32_t)(NpC + 16 * cc), v);
        tmem_ld_wait();
        if (DBG && dbg && blockIdx.x == 0 && j == 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) dbg[(2 * kTcRows + crow) * kTcDbgCols + 16 * cc + i] = __uint_as_float(v[i]);
        }
        if (smooth) {
          float term[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int o = 16 * cc + i;  // < Oy: tables are Oy long; padded entries are zero / masked
            const float x = fmaf(__uint_as_float(v[i]), t_P[o], t_Q[o]);  // (s - goal) * w, s = (y_raw + b3)*sd + mu
            term[i] = (fast_sqrt(fmaf(x, x, m.alpha2)) - m.alpha) * t_M[o];
          }
          st_total += ((term[0] + term[1]) + (term[2] + term[3])) + ((term[4] + term[5]) + (term[6] + term[7])) +
                      (((term[8] + term[9]) + (term[10] + term[11])) + ((term[12] + term[13]) + (term[14] + term[15])));
        } else {
          // dm_control task cost: pick the (at most four) un-normalised state entries it reads
          int pick[4];
          task_pick_indices(m.cost_kind, pick);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int o = 16 * cc + i;
            const float sv = fmaf(__uint_as_float(v[i]) + t_b3[o], t_sd[o], t_mu[o]);
#pragma unroll
            for (int q = 0; q < 4; ++q) p4[q] = (o == pick[q]) ? sv : p4[q];
          }
        }
        if (sout) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int o = 16 * cc + i;
            // unnormalize_state: y * std + mean   (data.py:255-257)
            if (o < O) sout[o] = fmaf(__uint_as_float(v[i]) + t_b3[o], t_sd[o], t_mu[o]);
          }
        }
      }
      if (!smooth) {
        // step j-1's control terms: written by the sampler before the action MMAs of that step were issued
        const int xs_slot = (j - 1) % kTcfXchSlots;
        const float a0 = xch ? xch[(xs_slot * 2 + 0) * kTcRows + crow] : 0.f, ctl = xch ? xch[(xs_slot * 2 + 1) * kTcRows + crow] : 0.f;
        st_total += task_cost(m.cost_kind, p4, a0, ctl);
      }
      if (j < H) {
        tc_fence_before();  // our tcgen05.ld of y is ordered before GEMM-A(j+1) overwrites it
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (crow == 0) { mbar_arrive(bar_y); if (DBG) tc_stamp(dbg, j - 1, 15); }  // completion #(j-1)
      }
    }
    costp[crow] = valid ? st_total : 0.f;
  } else {
    // ================= hidden-epilogue threads =================
    const int wg = warp >> 2, quarter = warp & 3;
    const int trow = quarter * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    for (int h = 0; h < H; ++h) {
      const uint32_t ph = h & 1;
#pragma unroll 1
      for (int layer = 0; layer < 2; ++layer) {
        const uint32_t dcol = layer == 0 ? 0u : (uint32_t)kTcD2Col;
        if (g.exp & 1) mbar_wait_nohint(layer == 0 ? bar_dA : bar_dB, ph); else mbar_wait(layer == 0 ? bar_dA : bar_dB, ph);
        tc_fence_after();
        if (tid == 0) if (DBG) tc_stamp(dbg, h, 4 + 2 * layer);
        const uint32_t bar_rel = layer == 0 ? bar_hA : bar_hB;
        // This warp's K-steps ks = wg, wg+4, wg+8, wg+12 (release groups 0..3), two at a time through two
        // register buffers: while one K-step is converted and stored, the TMEM load of the one after next
        // is already in flight, and a pair shares one wait::st / fence / arrive sequence.  (One K-step at a
        // time -- ld, wait, cvt, st, wait, fence, arrive: ~420 cycles each -- paced the MMAs at 1,690 cycles
        // per GEMM against 1,500 of MMA time; measured with profiles/tc_timeline.py.)
        auto convert_store = [&](const uint32_t (&v)[16], int ks) {
          if (DBG && dbg && blockIdx.x == 0 && h == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dbg[(layer * kTcRows + trow) * kTcDbgCols + 16 * ks + i] = __uint_as_float(v[i]);
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[i] = pack_relu<FP16>(v[2 * i], v[2 * i + 1]);
          tmem_st8(lane_base + dcol + 16 * ks, pk);  // 8 packed words in the first half of the K-step's own columns
        };
        uint32_t va[16], vb[16];
        const int k0 = wg, k1 = wg + 4, k2 = wg + 8, k3 = wg + 12;
        if (k0 < KS_H) tmem_ld16(lane_base + dcol + 16 * k0, va);
        if (k1 < KS_H) tmem_ld16(lane_base + dcol + 16 * k1, vb);
        tmem_ld_wait();
        if (k0 < KS_H) convert_store(va, k0);
        if (k2 < KS_H) tmem_ld16(lane_base + dcol + 16 * k2, va);
        if (!(g.exp & 4)) {  // release the first K-step group on its own, as early as possible (96.9 -> 93.8 us)
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0 && k0 < KS_H) mbar_arrive(bar_rel);
        }
        if (k1 < KS_H) convert_store(vb, k1);
        if (k3 < KS_H) tmem_ld16(lane_base + dcol + 16 * k3, vb);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if ((g.exp & 4) && k0 < KS_H) mbar_arrive(bar_rel);
          if (k1 < KS_H) mbar_arrive(bar_rel + 8);
        }
        if (DBG && lane == 0) {  // first pair of this warp: own stamp for warps 0/15, latest over all 16
          if (warp == 0 || warp == 15) tc_stamp(dbg, h, (warp == 0 ? 19 : 21) + layer);
          if (dbg && blockIdx.x == 1 && h < kTcTimelineSteps)
            atomicMax(reinterpret_cast<long long*>(dbg + kTcDbgFloats) + h * kTcTimelineEvents + 23 + layer, clock64());
        }
        if (k2 < KS_H) {
          tmem_ld_wait();
          convert_store(va, k2);
          if (k3 < KS_H) convert_store(vb, k3);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar_rel + 16);
            if (k3 < KS_H) mbar_arrive(bar_rel + 24);
          }
        }
        if (tid == 0) if (DBG) tc_stamp(dbg, h, 5 + 2 * layer);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid < kTcRows) {
    const long long row = (long long)blockIdx.x * kTcRows + tid;
    if (row < R) costs[row] = costp[tid] + costp[kTcRows + tid];
  }
  if (g.stamps && tid == 0) g.stamps[3 * blockIdx.x + 2] = (long long)globaltimer_ns();
  if (warp == kTcfMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// ---- host side of the tensor-core engines ---------------------------------------------------
enum TcKind { kTcFused = 0, kTcUnfused = 1, kTcWide = 2 };

struct TcModel {
  int ready = 0;
  bool fp16 = true;
  int kind = kTcFused;  // rollout_tcf_kernel; rollout_tc_kernel (unfused fallback); rollout_tcw_kernel (hidden > 255)
  TcGeom g{};
  TcfGeom fg{};
  TcwGeom wg{};
  bool head = false;  // weight-streaming kernel with the reward-head region (RewardAgent's cost)
  int w_bytes = 0;
  uint8_t* d_wimg = nullptr;
  float* d_dbg = nullptr;  // optional accumulator dump + timeline (tests / profiling)
};

inline bool tc_init(TcModel* t, int O, int A, int U, bool fp16, size_t max_smem, std::string* why, bool head = false) {
  t->fp16 = fp16;
  t->head = head;
  const char* force = getenv("MBRL_TC_UNFUSED");
  const char* wide = getenv("MBRL_TC_WIDE");  // tests: run small shapes through the streaming kernel
  std::string why_f, why_u;
  int stage_bytes = 0;  // geometry default
  if (const char* sk = getenv("MBRL_TCW_STAGE_KB")) { const int v = std::atoi(sk); if (v == 16 || v == 32) stage_bytes = v * 1024; }
  if ((wide && wide[0] == '1') || head) {  // the reward head is built into the weight-streaming kernel only
    if (!tcw_geometry(O, A, U, max_smem, &t->wg, why, stage_bytes, head)) return false;
    t->kind = kTcWide;
  } else if (!(force && force[0] == '1') && tcf_geometry(O, A, U, max_smem, &t->fg, &why_f)) {
    t->kind = kTcFused;
  } else if (tc_geometry(O, A, U, max_smem, &t->g, &why_u)) {
    t->kind = kTcUnfused;
  } else if (tcw_geometry(O, A, U, max_smem, &t->wg, why, stage_bytes)) {
    t->kind = kTcWide;
  } else {
    if (!why_u.empty()) *why += "; resident-weight kernel: " + why_u;
    if (!why_f.empty()) *why += "; fused: " + why_f;
    return false;
  }
  if (t->kind == kTcWide) {
    // CTAs per cluster sharing one multicast weight stream (MBRL_TCW_CLUSTER=1|2|4 overrides)
    t->wg.cluster = kTcwDefaultCluster;
    if (const char* x = getenv("MBRL_TCW_EXP")) t->wg.exp = std::atoi(x);
    if (const char* c = getenv("MBRL_TCW_CLUSTER")) {
      const int v = std::atoi(c);
      if (v == 1 || v == 2 || v == 4) t->wg.cluster = v;
    }
  }
  if (t->kind == kTcFused) { if (const char* x = getenv("MBRL_TCF_EXP")) t->fg.exp = std::atoi(x); }
  t->w_bytes = t->kind == kTcFused ? t->fg.w_bytes : (t->kind == kTcUnfused ? t->g.w_bytes : t->wg.w_bytes);
  if (cudaMalloc((void**)&t->d_wimg, t->w_bytes) != cudaSuccess) { *why = "cudaMalloc failed"; return false; }
  t->ready = 1;
  return true;
}

// dm_control task-cost epilogues (cost threads pick the entries, sampler threads hand over the
// control terms): built into the wide and the fused kernel (the fused one needs shared-memory room for
// the control-term slots when the cost depends on the control).
inline bool tc_supports_task_cost(const TcModel* t, int kind) {
  if (t->kind == kTcWide) return true;
  if (t->kind != kTcFused) return false;
  const bool needs_control = kind == MBRL_COST_DMC_CARTPOLE_SWINGUP || kind == MBRL_COST_DMC_HUMANOID_RUN;
  return !needs_control || t->fg.xch_off >= 0;
}

inline void tc_free(TcModel* t) {
  if (t->d_wimg) cudaFree(t->d_wimg);
  if (t->d_dbg) cudaFree(t->d_dbg);
  t->d_wimg = nullptr; t->d_dbg = nullptr; t->ready = 0;
}

inline bool tc_set_weights(TcModel* t, const float* W1, const float* b1, const float* W2, const float* b2,
                           const float* W3, const float* b3, std::string* why) {
  std::vector<uint16_t> img;
  if (t->kind == kTcFused) tcf_pack(t->fg, t->fp16, W1, b1, W2, b2, W3, b3, &img);
  else if (t->kind == kTcUnfused) tc_pack(t->g, t->fp16, W1, b1, W2, b2, W3, &img);  // b3 is added in fp32 in the last epilogue
  else tcw_pack(t->wg, t->fp16, W1, b1, W2, b2, W3, &img);
  if (cudaMemcpy(t->d_wimg, img.data(), t->w_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
    *why = "cudaMemcpy of packed operands failed";
    return false;
  }
  return true;
}

template <class Kern, class Geom>
inline cudaError_t tc_launch_one(Kern kern, const Geom& g, int smem_bytes, int threads, TcModel* t, const ModelDev& m,
                                 const ActionSource& src, const Shape& sh, const float* d_s0, float* d_costs,
                                 float* d_states, float* d_actions, cudaStream_t st, int cluster = 1) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)((sh.rows() + kTcRows - 1) / kTcRows);
  // programmatic dependent launch: the prologue (barriers, TMEM, weight TMA) overlaps the predecessor
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = (size_t)smem_bytes; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (!getenv("MBRL_NO_PDL")) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster > 1) {  // weight-streaming kernel: CTAs of a cluster share one multicast weight stream
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cluster; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
    cfg.gridDim = dim3((grid + cluster - 1) / cluster * cluster);  // surplus CTAs own no valid row but keep the ring protocol
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  const uint8_t* wimg = t->d_wimg;
  float* dbg = t->d_dbg;
  return cudaLaunchKernelEx(&cfg, kern, g, wimg, m, src, sh, d_s0, d_costs, d_states, d_actions, dbg);
}

inline cudaError_t tc_launch_rollout(TcModel* t, const ModelDev& m, const ActionSource& src, const Shape& sh,
                                     const float* d_s0, float* d_costs, float* d_states, float* d_actions, int num_sms,
                                     cudaStream_t st) {
  (void)num_sms;
  if (!t->ready) return cudaErrorNotReady;
  const bool dbg = t->d_dbg != nullptr;
#define MBRL_TC_LAUNCH(KERN, GEOM, THREADS) \
  return tc_launch_one(KERN, GEOM, (GEOM).smem_bytes, THREADS, t, m, src, sh, d_s0, d_costs, d_states, d_actions, st, \
                       t->kind == kTcWide ? t->wg.cluster : 1)
  if (t->kind == kTcFused) {
    const int spec = getenv("MBRL_TCF_NO_SPEC") ? 0 : tcf_matches_spec(t->fg);
    if (m.cost_kind != MBRL_COST_SMOOTHABS_COSH) {  // dm_control task cost (no debug-dump variant)
      if (dbg) return cudaErrorNotSupported;
      if (spec == 1 && t->fp16) MBRL_TC_LAUNCH((rollout_tcf_kernel<true, false, 1, true>), t->fg, kTcfThreads);
      if (spec == 1) MBRL_TC_LAUNCH((rollout_tcf_kernel<false, false, 1, true>), t->fg, kTcfThreads);
      if (spec == 2 && t->fp16) MBRL_TC_LAUNCH((rollout_tcf_kernel<true, false, 2, true>), t->fg, kTcfThreads);
      if (spec == 2) MBRL_TC_LAUNCH((rollout_tcf_kernel<false, false, 2, true>), t->fg, kTcfThreads);
      if (t->fp16) MBRL_TC_LAUNCH((rollout_tcf_kernel<true, false, 0, true>), t->fg, kTcfThreads);
      MBRL_TC_LAUNCH((rollout_tcf_kernel<false, false, 0, true>), t->fg, kTcfThreads);
    }
    if (spec == 1) {
      if (t->fp16 && !dbg) MBRL_TC_LAUNCH((rollout_tcf_kernel<true, false, 1, false>), t->fg, kTcfThreads);
      if (t->fp16) MBRL_TC_LAUNCH((rollout_tcf_kernel<true, true, 1, false>), t->fg, kTcfThreads);
      if (!dbg) MBRL_TC_LAUNCH((rollout_tcf_kernel<false, false, 1, false>), t->fg, kTcfThreads);
      MBRL_TC_LAUNCH((rollout_tcf_kernel<false, true, 1, false>), t->fg, kTcfThreads);
    }
    if (spec == 2 && !dbg) {  // (the debug dump of this class runs on the run-time-geometry instantiation)
      if (t->fp16) MBRL_TC_LAUNCH((rollout_tcf_kernel<true, false, 2, false>), t->fg, kTcfThreads);
      MBRL_TC_LAUNCH((rollout_tcf_kernel<false, false, 2, false>), t->fg, kTcfThreads);
    }
    if (t->fp16 && !dbg) MBRL_TC_LAUNCH((rollout_tcf_kernel<true, false, 0, false>), t->fg, kTcfThreads);
    if (t->fp16) MBRL_TC_LAUNCH((rollout_tcf_kernel<true, true, 0, false>), t->fg, kTcfThreads);
    if (!dbg) MBRL_TC_LAUNCH((rollout_tcf_kernel<false, false, 0, false>), t->fg, kTcfThreads);
    MBRL_TC_LAUNCH((rollout_tcf_kernel<false, true, 0, false>), t->fg, kTcfThreads);
  }
  if (t->kind == kTcWide) {
    const bool spec = tcw_matches_spec(t->wg) && !getenv("MBRL_TCW_NO_SPEC");
    if (spec) {
      if (t->fp16 && !dbg) MBRL_TC_LAUNCH((rollout_tcw_kernel<true, false, true>), t->wg, kTcwThreads);
      if (t->fp16) MBRL_TC_LAUNCH((rollout_tcw_kernel<true, true, true>), t->wg, kTcwThreads);
      if (!dbg) MBRL_TC_LAUNCH((rollout_tcw_kernel<false, false, true>), t->wg, kTcwThreads);
      MBRL_TC_LAUNCH((rollout_tcw_kernel<false, true, true>), t->wg, kTcwThreads);
    }
    if (t->fp16 && !dbg) MBRL_TC_LAUNCH((rollout_tcw_kernel<true, false, false>), t->wg, kTcwThreads);
    if (t->fp16) MBRL_TC_LAUNCH((rollout_tcw_kernel<true, true, false>), t->wg, kTcwThreads);
    if (!dbg) MBRL_TC_LAUNCH((rollout_tcw_kernel<false, false, false>), t->wg, kTcwThreads);
    MBRL_TC_LAUNCH((rollout_tcw_kernel<false, true, false>), t->wg, kTcwThreads);
  }
  if (t->fp16) MBRL_TC_LAUNCH(rollout_tc_kernel<true>, t->g, kTcThreads);
  MBRL_TC_LAUNCH(rollout_tc_kernel<false>, t->g, kTcThreads);
#undef MBRL_TC_LAUNCH
}

}  // namespace mbrl
