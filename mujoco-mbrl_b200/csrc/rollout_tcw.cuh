// Wide-hidden tcgen05 rollout engine: 256 < hidden <= 512 (BASELINE cfg 5: humanoid-run, O=67, A=21,
// U=512), where neither the weights (688 KB as 16-bit operands) fit shared memory nor a whole
// layer's fp32 accumulator plus its 16-bit input fit the 512 TMEM columns.
//
// Same computation as the other engines -- the hot loop of
// RandomShootingPlanner._generate_trajectories (src/mbrl/planners.py:199-210) with
// DynamicsModel.forward (src/mbrl/models.py:13-29), Model._forward (models.py:106-110), the
// normalisers (src/mbrl/data.py:255-260) and the per-step cost (models.py:244-272 or a dm_control task
// cost) -- one CTA per tile of 128 candidate rows for all H steps, nothing but costs[R] leaving the SM.
//
//   * WEIGHTS ARE STREAMED.  The host packs the operands ONCE in the exact order the MMAs consume
//     them (five parts per step: W1 rows [0,Nc) | W1 rows [Nc,Np) | W2 rows [0,Nc) + b2 K-step |
//     W2 rows [Nc,Np) + b2 | W3), each part a UMMA canonical K-major matrix whose K-step tiles are
//     contiguous.  A producer warp walks that image once per step and TMA-bulk-copies
//     (cp.async.bulk, mbarrier complete_tx) 16 KB stages into a shared-memory ring; the MMA warp
//     consumes a stage and hands it back with tcgen05.commit.  The image (688 KB) lives in L2.
//   * TMEM: columns [0,256) hold 16-bit activations (h1 packed two per column; later h2's upper half
//     and the layer-3 accumulator y at [256-Op, 256)), columns [256, 256+Nc) one fp32 accumulator for
//     a CHUNK of Nc = Np/2 hidden units.  Every hidden layer is two chunks.
//   * per step:  L1 chunk c: SS MMAs (A = the 128 x Kx input tile in shared memory, [actions | 1 |
//     state]) -> ACC; epilogue: relu -> 16 bit -> TMEM columns [c*Nc/2, ...).  L2 chunk 0: TS MMAs
//     (A = h1 from TMEM) + one SS K-step that adds b2 (A = a constant tile [1,0,..]) -> ACC; epilogue
//     -> h2's lower half as a canonical A tile in SHARED memory (h1 is still live in TMEM).  L2
//     chunk 1 -> ACC; epilogue -> TMEM columns [0, Nc/2) (h1 is dead).  L3: SS K-steps over the lower
//     half + TS K-steps over the upper half (released per epilogue round) -> y.  Cost warps: y + b3
//     (fp32), un-normalise, cost, and the normalised prediction becomes the state section of the
//     next step's input tile.  Actions are sampled one step ahead by the sampler warps.
//
// Warp roles (832 threads, 1 CTA/SM): warps 0-15 hidden epilogues (warpgroup g owns the 32-column
// units u = g, g+4 of a chunk; warp%4 = TMEM lane quarter), 16-19 cost (one thread per row), 20-23
// sampler (one thread per row), 24 MMA issue (one elected lane), 25 weight producer.
#pragma once
#include "rollout_tc.cuh"

namespace mbrl {

constexpr int kTcwEpiWarps = 16;
constexpr int kTcwCostWarp0 = kTcwEpiWarps;
constexpr int kTcwSampWarp0 = kTcwCostWarp0 + 4;
constexpr int kTcwMmaWarp = kTcwSampWarp0 + 4;
constexpr int kTcwTmaWarp = kTcwMmaWarp + 1;
constexpr int kTcwThreads = (kTcwTmaWarp + 1) * 32;
constexpr int kTcwAccCol = 256;      // TMEM column of the chunk accumulator
constexpr int kTcwMaxStages = 8;
constexpr int kTcwStageBytes = 16384;
constexpr int kTcwBarriers = 2 * kTcwMaxStages + 6;  // full[8], empty[8], x, acc, e0, e1, l1, y

struct TcwGeom {
  int O, A, U;
  int Ka;   // action section of the input tile: A actions, the constant 1, zero pad (multiple of 8)
  int Kx;   // layer-1 K (multiple of 16)
  int Np;   // padded hidden width (multiple of 64)
  int Nc;   // chunk width Np/2 (multiple of 32, <= 256)
  int Op;   // layer-3 N (multiple of 16)
  int KH;   // hidden K-steps Np/16
  int QA, SC;  // 8-wide chunks of the action / state section
  int tile_h, tile_y;  // bytes of one K-step tile of a hidden part / of W3
  int tps_h, tps_y;    // tiles per ring stage
  int p1_bytes, p2_bytes, p3_bytes, w_bytes;
  int stages, ycol;
  int tab_off, xa_off, xs_off, one_off, h2_off, ring_off, bar_off, ms_off, ms_floats, smem_bytes;
};

inline bool tcw_geometry(int O, int A, int U, size_t max_smem, TcwGeom* g, std::string* why) {
  g->O = O; g->A = A; g->U = U;
  g->Ka = round_up(A + 1, 8);
  g->Kx = round_up(g->Ka + O, 16);
  g->Np = round_up(U, 64);
  g->Nc = g->Np / 2;
  g->Op = round_up(O, 16);
  g->KH = g->Np / 16;
  if (g->Np > 512) { *why = "hidden > 512 unsupported by the tensor-core engines (TMEM holds 512 columns)"; return false; }
  if (g->Op > 128) { *why = "obs_dim > 128 unsupported"; return false; }
  g->QA = g->Ka / 8;
  g->SC = (g->Kx - g->Ka) / 8;
  g->tile_h = g->Nc * 32;
  g->tile_y = g->Op * 32;
  g->tps_h = kTcwStageBytes / g->tile_h;
  g->tps_y = kTcwStageBytes / g->tile_y;
  g->p1_bytes = (g->Kx / 16) * g->tile_h;
  g->p2_bytes = (g->KH + 1) * g->tile_h;
  g->p3_bytes = g->KH * g->tile_y;
  g->w_bytes = 2 * g->p1_bytes + 2 * g->p2_bytes + g->p3_bytes;
  g->ycol = 256 - g->Op;
  // fp32 tables: 6 of Op (b3, P, Q, sd, mu, mask), 2 of kMaxAct, cost partials [2][128], task-cost
  // exchange [2 step parities][a0, ctl][128]
  g->tab_off = 0;
  g->xa_off = round_up((6 * g->Op + 2 * kMaxAct + 2 * kTcRows + 4 * kTcRows) * 4, 128);
  g->xs_off = g->xa_off + 2 * g->QA * 2048;
  g->one_off = g->xs_off + g->SC * 2048;
  g->h2_off = g->one_off + 4096;
  g->ring_off = g->h2_off + g->Nc * 256;
  const long long fixed = (long long)g->ring_off + 8 * kTcwBarriers + 16;
  const long long room = (long long)max_smem - fixed;
  g->stages = (int)std::min<long long>(kTcwMaxStages, room / kTcwStageBytes);
  if (g->stages < 3) { *why = "not enough shared memory for the weight ring"; return false; }
  g->bar_off = g->ring_off + g->stages * kTcwStageBytes;
  g->smem_bytes = g->bar_off + 8 * kTcwBarriers + 16;
  g->ms_off = g->smem_bytes;
  g->ms_floats = (int)std::min<size_t>((max_smem - (size_t)g->smem_bytes) / 4, 4096);
  g->smem_bytes += 4 * g->ms_floats;
  return true;
}

// Operand image in consumption order (see the header).  Hidden unit u = c*Nc + n is row n of
// chunk c; rows / K entries beyond U stay zero, so padded units are relu(0) = 0.
inline void tcw_pack(const TcwGeom& g, bool fp16, const float* W1, const float* b1, const float* W2, const float* b2,
                     const float* W3, std::vector<uint16_t>* out) {
  const int O = g.O, A = g.A, U = g.U, D = O + A;
  std::vector<uint16_t>& img = *out;
  img.assign((size_t)g.w_bytes / 2, 0);
  for (int c = 0; c < 2; ++c) {
    const int off1 = c * g.p1_bytes, off2 = 2 * g.p1_bytes + c * g.p2_bytes;
    for (int n = 0; n < g.Nc; ++n) {
      const int u = c * g.Nc + n;
      if (u >= U) break;
      // layer 1: input column e:  e < A -> action e;  e == A -> constant 1 (carries b1);  Ka <= e -> state e-Ka
      for (int a = 0; a < A; ++a) tc_put(img, off1, g.Nc, n, a, W1[(size_t)u * D + O + a], fp16);
      tc_put(img, off1, g.Nc, n, A, b1[u], fp16);
      for (int o = 0; o < O; ++o) tc_put(img, off1, g.Nc, n, g.Ka + o, W1[(size_t)u * D + o], fp16);
      // layer 2: K = hidden units, then one extra K-step whose first entry multiplies the constant tile
      for (int k = 0; k < U; ++k) tc_put(img, off2, g.Nc, n, k, W2[(size_t)u * U + k], fp16);
      tc_put(img, off2, g.Nc, n, g.Np, b2[u], fp16);
    }
  }
  const int off3 = 2 * g.p1_bytes + 2 * g.p2_bytes;
  for (int o = 0; o < O; ++o)
    for (int k = 0; k < U; ++k) tc_put(img, off3, g.Op, o, k, W3[(size_t)o * U + k], fp16);
}

template <bool FP16, bool DBG>
__global__ void __launch_bounds__(kTcwThreads, 1)
rollout_tcw_kernel(TcwGeom g, const uint8_t* __restrict__ wimg, ModelDev m, ActionSource src, Shape sh,
                   const float* __restrict__ s0, float* __restrict__ costs, float* __restrict__ states_out,
                   float* __restrict__ actions_out, float* __restrict__ dbg) {
  extern __shared__ __align__(128) uint8_t tcw_smem[];
  uint8_t* const smem = tcw_smem;
  // PDL: everything before the sampler / cost threads' pdl_wait touches static model data only
  pdl_trigger();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int O = g.O, A = g.A, H = sh.H;
  const int KS_X = g.Kx >> 4, KH = g.KH, KC = g.Nc >> 4, NU = g.Nc >> 5;
  const int QA = g.QA, SC = g.SC, S = g.stages;
  const bool smooth = m.cost_kind == MBRL_COST_SMOOTHABS_COSH;

  float* tab = reinterpret_cast<float*>(smem + g.tab_off);
  float *t_b3 = tab, *t_P = tab + g.Op, *t_Q = tab + 2 * g.Op, *t_sd = tab + 3 * g.Op, *t_mu = tab + 4 * g.Op;
  float *t_M = tab + 5 * g.Op;
  float *t_ainv = tab + 6 * g.Op, *t_aoff = t_ainv + kMaxAct, *costp = t_aoff + kMaxAct;
  float* xch = costp + 2 * kTcRows;  // [parity][0: a0, 1: ctl_mean][row]
  uint8_t* const xa = smem + g.xa_off;
  uint8_t* const xs = smem + g.xs_off;
  uint8_t* const h2lo = smem + g.h2_off;
  const int xa_bytes = QA * 2048;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.bar_off);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + g.bar_off + 8 * kTcwBarriers);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kTcwMaxStages;
  const uint32_t bar_x = bar_empty + 8 * kTcwMaxStages, bar_acc = bar_x + 8, bar_e0 = bar_x + 16, bar_e1 = bar_x + 24;
  const uint32_t bar_l1 = bar_x + 32, bar_y = bar_x + 40;

  if (warp == kTcwMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < kTcwMaxStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
      mbar_init(bar_x, 2);  // sampler group + cost group
      mbar_init(bar_acc, 1); mbar_init(bar_e0, kTcwEpiWarps); mbar_init(bar_e1, kTcwEpiWarps);
      mbar_init(bar_l1, 1); mbar_init(bar_y, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < g.Op; i += kTcwThreads) {
    const bool in = i < O;
    t_M[i] = in ? 1.f : 0.f;
    const float b3 = in ? __ldg(m.b3 + i) : 0.f, sd = in ? __ldg(m.sd_s + i) : 1.f, mu = in ? __ldg(m.mu_s + i) : 0.f;
    const float w = (in && smooth) ? __ldg(m.cost_w + i) : 0.f, goal = (in && smooth) ? __ldg(m.goal + i) : 0.f;
    t_b3[i] = b3; t_sd[i] = sd; t_mu[i] = mu;
    t_P[i] = sd * w;
    t_Q[i] = (b3 * sd + mu - goal) * w;
  }
  for (int i = tid; i < kMaxAct; i += kTcwThreads) {
    // normalised action = a * inv - off; the constant-1 column (i == A) is 0 * 0 - (-1)
    const float inv = i < A ? 1.0f / __ldg(m.sd_a + i) : 0.f;
    t_ainv[i] = inv;
    t_aoff[i] = i < A ? __ldg(m.mu_a + i) * inv : (i == A ? -1.f : 0.f);
  }
  // the constant A tile of the b2 K-step: element (row, k = 0) = 1, everything else 0
  for (int i = tid; i < 256; i += kTcwThreads) {
    const uint32_t one = FP16 ? 0x3C00u : 0x3F80u;
    reinterpret_cast<uint4*>(smem + g.one_off)[i] = make_uint4(i < kTcRows ? one : 0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const long long R = sh.rows();

  if (warp == kTcwTmaWarp) {
    // ================= weight producer =================
    if (elect_one()) {
      const uint32_t ring = smem_u32(smem + g.ring_off);
      uint32_t st = 0, ph = 0;
      for (int h = 0; h < H; ++h) {
        uint32_t off = 0;
#pragma unroll 1
        for (int part = 0; part < 5; ++part) {
          const uint32_t total = (uint32_t)(part < 2 ? g.p1_bytes : (part < 4 ? g.p2_bytes : g.p3_bytes));
          const uint32_t per = (uint32_t)(part < 4 ? g.tps_h * g.tile_h : g.tps_y * g.tile_y);
#pragma unroll 1
          for (uint32_t done = 0; done < total; done += per) {
            const uint32_t n = min(per, total - done);
            mbar_wait(bar_empty + 8 * st, ph ^ 1);  // first lap: passes at once on a fresh barrier
            mbar_arrive_expect_tx(bar_full + 8 * st, n);
            bulk_g2s(ring + st * kTcwStageBytes, wimg + off, n, bar_full + 8 * st);
            off += n;
            if (++st == (uint32_t)S) { st = 0; ph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kTcwMmaWarp) {
    // ================= MMA issuer warp (converged; one elected lane issues) =================
    const uint32_t idesc_h = umma_idesc(g.Nc, FP16), idesc_y = umma_idesc(g.Op, FP16);
    const uint32_t ring = smem_u32(smem + g.ring_off);
    const uint64_t d_ring_h = umma_desc(ring, (uint32_t)g.Nc * 16, 128);
    const uint64_t d_ring_y = umma_desc(ring, (uint32_t)g.Op * 16, 128);
    const uint64_t d_one = umma_desc(smem_u32(smem + g.one_off), 2048, 128);
    const uint64_t d_h2 = umma_desc(smem_u32(h2lo), 2048, 128);
    const uint32_t xa0 = smem_u32(xa), xs0 = smem_u32(xs);
    const uint32_t tm_h = tmem, tm_acc = tmem + kTcwAccCol, tm_y = tmem + (uint32_t)g.ycol;
    uint32_t st = 0, ph = 0, e0_cnt = 0, e1_cnt = 0;
    auto wait_epi = [&](int r) {
      if (r == 0) { mbar_wait(bar_e0, e0_cnt & 1); ++e0_cnt; }
      else { mbar_wait(bar_e1, e1_cnt & 1); ++e1_cnt; }
      tc_fence_after();
    };
    auto advance = [&]() { if (++st == (uint32_t)S) { st = 0; ph ^= 1; } };

    for (int h = 0; h < H; ++h) {
      mbar_wait(bar_x, h & 1);
      tc_fence_after();
      const uint32_t xa_cur = xa0 + (uint32_t)((h & 1) * xa_bytes);
      // ---- layer 1, two chunks: ACC = x . W1[chunk]^T ----
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        if (c == 1) { wait_epi(0); wait_epi(1); }  // ACC drained by the epilogue of chunk 0
#pragma unroll 1
        for (int t0 = 0; t0 < KS_X; t0 += g.tps_h) {
          const int n = min(g.tps_h, KS_X - t0);
          mbar_wait(bar_full + 8 * st, ph);
          tc_fence_after();
          if (elect_one()) {
            for (int i = 0; i < n; ++i) {
              // K-step ks of the input tile = its 8-wide chunks 2ks, 2ks+1; the action chunks live in
              // the double-buffered action tile, the state chunks in the state tile: LBO is simply
              // the distance between the two chunks
              const int ks = t0 + i, c0 = 2 * ks, c1 = c0 + 1;
              const uint32_t a0 = c0 < QA ? xa_cur + (uint32_t)c0 * 2048u : xs0 + (uint32_t)(c0 - QA) * 2048u;
              const uint32_t a1 = c1 < QA ? xa_cur + (uint32_t)c1 * 2048u : xs0 + (uint32_t)(c1 - QA) * 2048u;
              const uint64_t bd = d_ring_h + (uint64_t)((st * kTcwStageBytes + (uint32_t)(i * g.tile_h)) >> 4);
              mma_ss(tm_acc, umma_desc(a0, a1 - a0, 128), bd, idesc_h, ks > 0);
            }
            tc_commit(bar_empty + 8 * st);
            if (t0 + n == KS_X) {
              tc_commit(bar_acc);
              if (c == 1) tc_commit(bar_l1);  // the action tile of this step is free again
            }
          }
          __syncwarp();
          advance();
        }
      }
      // ---- layer 2, two chunks: ACC = h1 . W2[chunk]^T + b2 ----
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        wait_epi(0); wait_epi(1);  // c == 0: h1 complete in TMEM; c == 1: h2's lower half in shared memory
#pragma unroll 1
        for (int t0 = 0; t0 <= KH; t0 += g.tps_h) {
          const int n = min(g.tps_h, KH + 1 - t0);
          mbar_wait(bar_full + 8 * st, ph);
          tc_fence_after();
          if (elect_one()) {
            for (int i = 0; i < n; ++i) {
              const int ks = t0 + i;
              const uint64_t bd = d_ring_h + (uint64_t)((st * kTcwStageBytes + (uint32_t)(i * g.tile_h)) >> 4);
              if (ks < KH) mma_ts(tm_acc, tm_h + 8u * (uint32_t)ks, bd, idesc_h, ks > 0);
              else mma_ss(tm_acc, d_one, bd, idesc_h, 1);
            }
            tc_commit(bar_empty + 8 * st);
            if (t0 + n == KH + 1) tc_commit(bar_acc);
          }
          __syncwarp();
          advance();
        }
      }
      // ---- layer 3: y = h2 . W3^T; lower half from shared memory, upper half from TMEM as released ----
      {
        bool w0 = false, w1 = false;
#pragma unroll 1
        for (int t0 = 0; t0 < KH; t0 += g.tps_y) {
          const int n = min(g.tps_y, KH - t0);
          mbar_wait(bar_full + 8 * st, ph);
          tc_fence_after();
          const int last_ks = t0 + n - 1;
          if (!w0 && last_ks >= KC) { wait_epi(0); w0 = true; }
          if (!w1 && last_ks >= KC + 8) { wait_epi(1); w1 = true; }
          if (elect_one()) {
            for (int i = 0; i < n; ++i) {
              const int ks = t0 + i;
              const uint64_t bd = d_ring_y + (uint64_t)((st * kTcwStageBytes + (uint32_t)(i * g.tile_y)) >> 4);
              if (ks < KC) mma_ss(tm_y, d_h2 + (uint64_t)(ks * (4096 >> 4)), bd, idesc_y, ks > 0);
              else mma_ts(tm_y, tm_h + 8u * (uint32_t)(ks - KC), bd, idesc_y, 1);
            }
            tc_commit(bar_empty + 8 * st);
            if (t0 + n == KH) tc_commit(bar_y);
          }
          __syncwarp();
          advance();
        }
        if (!w0) wait_epi(0);
        if (!w1) wait_epi(1);  // (a round without units still completes: keep the phases in step)
      }
    }
  } else if (warp >= kTcwSampWarp0) {
    // ================= sampler threads (one per row) =================
    const int srow = tid - kTcwSampWarp0 * 32;
    const long long row = (long long)blockIdx.x * kTcRows + srow;
    const bool valid = row < R;
    const int env_l = valid ? (int)(row / sh.N) : 0;
    const int cand_l = valid ? (int)(row - (long long)env_l * sh.N) : 0;
    const float inv_beta = 1.0f / m.beta, cscale = (valid && smooth) ? m.beta2 / (float)A : 0.f;
    const float inv_A = 1.0f / (float)A;
    float act_total = 0.f;
    pdl_wait();

    // stage the sampling mean/std rows of this tile's environments in shared memory (see rollout_tcf.cuh)
    const bool gauss = src.mode == MBRL_SAMPLE_INJECT_NOISE || src.mode == MBRL_SAMPLE_GAUSSIAN;
    const long long row_lo = (long long)blockIdx.x * kTcRows, row_hi = min(row_lo + kTcRows, R) - 1;
    const int env_lo = (int)(row_lo / sh.N), env_hi = (int)(row_hi / sh.N);
    const int ms_n = (env_hi - env_lo + 1) * H * A;
    const bool staged = gauss && 2 * ms_n <= g.ms_floats;
    float* const ms_mu = reinterpret_cast<float*>(smem + g.ms_off);
    float* const ms_sd = ms_mu + ms_n;
    if (staged) {
      const long long base = (long long)env_lo * H * A;
      for (int i = srow; i < ms_n; i += kTcRows) { ms_mu[i] = dep_load(src.mu + base + i); ms_sd[i] = dep_load(src.sd + base + i); }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    const int ms_env = (env_l - env_lo) * H * A;

    for (int hs = 0; hs < H; ++hs) {
      // layer 1 of step hs-1 has finished reading action tile (hs-1)&1; tile hs&1 was last read by
      // step hs-2, and the exchange slot hs&1 by the cost threads of step hs-2
      if (hs >= 1) mbar_wait(bar_l1, (hs - 1) & 1);
      float acc = 0.f, ctl = 0.f, first = 0.f;
      float* aout = (actions_out && valid) ? actions_out + ((long long)hs * R + row) * A : nullptr;
      uint8_t* xt = xa + (hs & 1) * xa_bytes;
      const float* pm = staged ? ms_mu + ms_env + hs * A : nullptr;
      const float* ps = staged ? ms_sd + ms_env + hs * A : nullptr;
      for (int q = 0; q < QA; ++q) {
        float v[8];
        {
          float t4[4], u4[4];
          if (8 * q < A) raw_action4(src, A, H, hs, env_l, cand_l, row, R, 2 * q, t4, pm, ps);
          else { t4[0] = t4[1] = t4[2] = t4[3] = 0.f; }
          if (8 * q + 4 < A) raw_action4(src, A, H, hs, env_l, cand_l, row, R, 2 * q + 1, u4, pm, ps);
          else { u4[0] = u4[1] = u4[2] = u4[3] = 0.f; }
          v[0] = t4[0]; v[1] = t4[1]; v[2] = t4[2]; v[3] = t4[3];
          v[4] = u4[0]; v[5] = u4[1]; v[6] = u4[2]; v[7] = u4[3];
        }
        if (q == 0) first = v[0];
        // branch-free: raw_action4 returns 0 beyond A (cosh(0) - 1 == 0), tables are zero-padded
        float xn[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc += cosh_m1_fast(v[i] * inv_beta);
          const float quad = fabsf(v[i]) < 1.0f ? 1.0f - v[i] * v[i] : 0.0f;
          ctl += (8 * q + i < A) ? quad : 0.0f;
          xn[i] = fmaf(v[i], t_ainv[8 * q + i], -t_aoff[8 * q + i]);
        }
        if (aout) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (8 * q + i < A) aout[8 * q + i] = v[i];
        }
        uint4 pk;
        pk.x = pack2<FP16>(xn[0], xn[1]); pk.y = pack2<FP16>(xn[2], xn[3]);
        pk.z = pack2<FP16>(xn[4], xn[5]); pk.w = pack2<FP16>(xn[6], xn[7]);
        *reinterpret_cast<uint4*>(xt + q * 2048 + srow * 16) = pk;
      }
      act_total = fmaf(cscale, acc, act_total);  // CoshLoss: beta^2 * mean_a(cosh(a/beta) - 1)
      xch[((hs & 1) * 2 + 0) * kTcRows + srow] = first;
      xch[((hs & 1) * 2 + 1) * kTcRows + srow] = ctl * inv_A;
      fence_proxy_async();   // generic-proxy tile writes -> visible to the MMA (async proxy)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (srow == 0) mbar_arrive(bar_x);
    }
    costp[kTcRows + srow] = act_total;
  } else if (warp >= kTcwCostWarp0) {
    // ================= cost threads (one per row) =================
    const int crow = tid - kTcwCostWarp0 * 32;
    const long long row = (long long)blockIdx.x * kTcRows + crow;
    const bool valid = row < R;
    const int env_l = valid ? (int)(row / sh.N) : 0;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    int pick[4];
    task_pick_indices(m.cost_kind, pick);
    float st_total = 0.f;
    pdl_wait();
    // step 0: the normalised initial state
    for (int j = 0; j < SC; ++j) {
      float xn[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int o = 8 * j + i;
        xn[i] = (o < O && valid) ? (dep_load(s0 + (long long)env_l * O + o) - t_mu[o]) / t_sd[o] : 0.f;
      }
      uint4 pk;
      pk.x = pack2<FP16>(xn[0], xn[1]); pk.y = pack2<FP16>(xn[2], xn[3]);
      pk.z = pack2<FP16>(xn[4], xn[5]); pk.w = pack2<FP16>(xn[6], xn[7]);
      *reinterpret_cast<uint4*>(xs + j * 2048 + crow * 16) = pk;
    }
    fence_proxy_async();
    asm volatile("bar.sync 2, 128;" ::: "memory");
    if (crow == 0) mbar_arrive(bar_x);

    for (int h = 0; h < H; ++h) {
      mbar_wait(bar_y, h & 1);
      tc_fence_after();
      float* sout = (states_out && valid) ? states_out + ((long long)h * R + row) * O : nullptr;
      float p4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int cc = 0; cc < (g.Op >> 4); ++cc) {
        uint32_t v[32];
        tmem_ld16(lane_base + (uint32_t)(g.ycol + 16 * cc), v);
        tmem_ld_wait();
        if (DBG && dbg && blockIdx.x == 0 && h == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) dbg[(2 * kTcRows + crow) * kTcDbgCols + 16 * cc + i] = __uint_as_float(v[i]);
        }
        float y[16], term[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int o = 16 * cc + i;  // < Op: tables are Op long, padded entries are zero / masked
          const float raw = __uint_as_float(v[i]);
          y[i] = (raw + t_b3[o]) * t_M[o];              // normalised prediction == next input
          const float x = fmaf(raw, t_P[o], t_Q[o]);    // (s - goal) * w with s = y*sd + mu
          term[i] = (fast_sqrt(fmaf(x, x, m.alpha2)) - m.alpha) * t_M[o];
        }
        if (smooth)
          st_total += ((term[0] + term[1]) + (term[2] + term[3])) + ((term[4] + term[5]) + (term[6] + term[7])) +
                      (((term[8] + term[9]) + (term[10] + term[11])) + ((term[12] + term[13]) + (term[14] + term[15])));
        else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int o = 16 * cc + i;
            const float s = fmaf(y[i], t_sd[o], t_mu[o]);
#pragma unroll
            for (int q = 0; q < 4; ++q) p4[q] = (o == pick[q]) ? s : p4[q];
          }
        }
        if (sout) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int o = 16 * cc + i;
            if (o < O) sout[o] = fmaf(y[i], t_sd[o], t_mu[o]);  // unnormalize_state (data.py:255-257)
          }
        }
        if (h + 1 < H) {
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int j = 2 * cc + jj;
            if (j < SC) {
              uint4 pk;
              pk.x = pack2<FP16>(y[8 * jj + 0], y[8 * jj + 1]); pk.y = pack2<FP16>(y[8 * jj + 2], y[8 * jj + 3]);
              pk.z = pack2<FP16>(y[8 * jj + 4], y[8 * jj + 5]); pk.w = pack2<FP16>(y[8 * jj + 6], y[8 * jj + 7]);
              *reinterpret_cast<uint4*>(xs + j * 2048 + crow * 16) = pk;
            }
          }
        }
      }
      if (!smooth)
        st_total += task_cost(m.cost_kind, p4, xch[((h & 1) * 2 + 0) * kTcRows + crow], xch[((h & 1) * 2 + 1) * kTcRows + crow]);
      if (h + 1 < H) {
        fence_proxy_async();   // generic-proxy writes of the state tile -> visible to the MMA
        tc_fence_before();     // our tcgen05.ld of y is ordered before the columns are reused
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (crow == 0) mbar_arrive(bar_x);
      }
    }
    costp[crow] = valid ? st_total : 0.f;
  } else {
    // ================= hidden-epilogue threads =================
    const int wg = warp >> 2, quarter = warp & 3;
    const int trow = quarter * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    uint32_t acc_cnt = 0;
    for (int h = 0; h < H; ++h) {
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {  // L1 chunk 0, L1 chunk 1, L2 chunk 0, L2 chunk 1
        mbar_wait(bar_acc, acc_cnt & 1);
        ++acc_cnt;
        tc_fence_after();
#pragma unroll 1
        for (int r = 0; r < 2; ++r) {
          const int u = 4 * r + wg;  // 32-column unit of the chunk
          if (u < NU) {
            uint32_t v[32], pk[16];
            tmem_ld32(lane_base + (uint32_t)(kTcwAccCol + 32 * u), v);
            tmem_ld_wait();
            if (DBG && dbg && blockIdx.x == 0 && h == 0) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                dbg[((q >> 1) * kTcRows + trow) * kTcDbgCols + (q & 1) * g.Nc + 32 * u + i] = __uint_as_float(v[i]);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack_relu<FP16>(v[2 * i], v[2 * i + 1]);
            if (q == 2) {
              // h2's lower half: canonical A tile in shared memory, 8 hidden units per 16-byte row chunk
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(h2lo + (4 * u + j) * 2048 + trow * 16) =
                    make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
              fence_proxy_async();
            } else {
              tmem_st16(lane_base + (uint32_t)((q == 1 ? g.Nc >> 1 : 0) + 16 * u), pk);
              tmem_st_wait();
            }
            tc_fence_before();
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(r == 0 ? bar_e0 : bar_e1);
        }
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
  if (warp == kTcwMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace mbrl
