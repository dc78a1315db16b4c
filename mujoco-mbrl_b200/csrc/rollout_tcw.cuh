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
constexpr int kTcwStageBytes = 16384;  // default ring stage (MBRL_TCW_STAGE_KB overrides: 16 or 32)
constexpr int kTcwDefaultCluster = 1;  // CTAs per cluster sharing one weight stream (see tc_init)
constexpr int kTcwBarriers = 2 * kTcwMaxStages + 6;  // full[8], empty[8], x, acc, e0, e1, l1, y
constexpr int kTcwMaxKx = 16;  // layer-1 K-steps (Kx <= 32 + 128 + 15)

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
  int cluster;  // CTAs per thread-block cluster sharing one weight stream (1, 2 or 4)
  int stage_bytes;  // ring stage size
  int exp;      // profiling experiments (MBRL_TCW_EXP bit mask; results are garbage when non-zero)
  int tab_off, xa_off, xs_off, one_off, h2_off, ring_off, bar_off, ms_off, ms_floats, smem_bytes;
  int head_off;  // reward-head cost: fp32 W4 [Np] + per-row partial sums [16 slots][128]; -1 when not enabled
};

constexpr int kTcwHeadSlots = 16;  // reward-head partial sums: [layer-2 chunk][32-column unit]

inline bool tcw_geometry(int O, int A, int U, size_t max_smem, TcwGeom* g, std::string* why, int stage_bytes = 0, bool head = false) {
  g->O = O; g->A = A; g->U = U;
  // A ring stage should hold >= 4 hidden K-step tiles: the MMA thread's per-stage bookkeeping (~150
  // cycles) then hides behind the MMAs it has queued (measured: 16 KB stages 11.4 ms, 32 KB 9.2 ms at cfg 5).
  if (stage_bytes == 0) stage_bytes = round_up(U, 64) / 2 * 32 * 4 > kTcwStageBytes ? 2 * kTcwStageBytes : kTcwStageBytes;
  g->stage_bytes = stage_bytes;
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
  g->tps_h = stage_bytes / g->tile_h;
  g->tps_y = stage_bytes / g->tile_y;
  g->p1_bytes = (g->Kx / 16) * g->tile_h;
  g->p2_bytes = (g->KH + 1) * g->tile_h;
  g->p3_bytes = g->KH * g->tile_y;
  g->w_bytes = 2 * g->p1_bytes + 2 * g->p2_bytes + g->p3_bytes;
  g->ycol = 256 - g->Op;
  g->cluster = 1;
  g->exp = 0;
  // fp32 tables: 6 of Op (b3, P, Q, sd, mu, mask), 2 of kMaxAct, cost partials [6][128] (five column
  // shares of the state cost + the action cost), task-cost exchange [2 step parities][a0, ctl][128]
  // and picked state entries [2 step parities][4][128]
  g->tab_off = 0;
  g->xa_off = round_up((6 * g->Op + 2 * kMaxAct + 6 * kTcRows + 4 * kTcRows + 8 * kTcRows) * 4, 128);
  g->xs_off = g->xa_off + 2 * g->QA * 2048;
  g->one_off = g->xs_off + g->SC * 2048;
  g->h2_off = g->one_off + 4096;
  g->head_off = -1;
  int after_h2 = g->h2_off + g->Nc * 256;
  if (head) { g->head_off = after_h2; after_h2 += round_up((g->Np + kTcwHeadSlots * kTcRows) * 4, 128); }
  g->ring_off = after_h2;
  const long long fixed = (long long)g->ring_off + 8 * kTcwBarriers + 16 + 8 * kTcwMaxKx;
  const long long room = (long long)max_smem - fixed;
  g->stages = (int)std::min<long long>(kTcwMaxStages, room / stage_bytes);
  if (g->stages < 3) { *why = "not enough shared memory for the weight ring"; return false; }
  g->bar_off = g->ring_off + g->stages * stage_bytes;
  g->smem_bytes = g->bar_off + 8 * kTcwBarriers + 16 + 8 * kTcwMaxKx;  // barriers, TMEM slot, layer-1 A-descriptor table
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

// ---- thread-block-cluster variants of the ring primitives ------------------------------------
// With a cluster of C CTAs every CTA copies 1/C of each stage and MULTICASTS it into the same
// ring slot of all C CTAs (the data and the mbarrier complete_tx land at the same CTA-relative
// offsets in every destination), so the stream is read from L2 once per cluster instead of once per
// CTA; a slot is handed back to all C producers at once by a multicast tcgen05.commit.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

// One 16-column half of a hidden-epilogue unit: fp32 accumulators -> relu -> 16 bit -> 8 packed words,
// stored to TMEM (mode 0: the next layer's A operand) or to h2's shared-memory A tile (mode 1: two 16-byte
// row chunks of 8 hidden units each); mode 2 is the reward head (ModelWithReward.linear4,
// src/mbrl/models.py:135-141): relu in fp32 and a dot product with 16 entries of W4 instead of a store.
template <bool FP16, bool DBG>
__device__ __forceinline__ void tcw_epi_half(const uint32_t (&v)[16], int mode, uint32_t tmem_dst, uint8_t* smem_dst, float* dbg,
                                             int h, int trow, int dbg_layer, int dbg_col, const float* w4, float& head_acc) {
  if (DBG && dbg && blockIdx.x == 0 && h == 0 && dbg_layer >= 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dbg[(dbg_layer * kTcRows + trow) * kTcDbgCols + dbg_col + i] = __uint_as_float(v[i]);
  }
  if (mode == 2) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 w = *reinterpret_cast<const float4*>(w4 + i);
      acc = fmaf(fmaxf(__uint_as_float(v[i + 0]), 0.f), w.x, acc);
      acc = fmaf(fmaxf(__uint_as_float(v[i + 1]), 0.f), w.y, acc);
      acc = fmaf(fmaxf(__uint_as_float(v[i + 2]), 0.f), w.z, acc);
      acc = fmaf(fmaxf(__uint_as_float(v[i + 3]), 0.f), w.w, acc);
    }
    head_acc += acc;
    return;
  }
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) pk[i] = pack_relu<FP16>(v[2 * i], v[2 * i + 1]);
  if (mode == 1) {
    *reinterpret_cast<uint4*>(smem_dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(smem_dst + 2048) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  } else {
    tmem_st8(tmem_dst, pk);
  }
}

// non-blocking probe of an mbarrier phase
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}

// SPEC: the geometry class of BASELINE cfg 5 (humanoid-run: Kx = 96, hidden 449..512, Op = 80, 32 KB ring
// stages, 3 stages) as compile-time constants.  What limits this kernel is the dependent-instruction
// chain of the ONE thread that issues the MMAs; with constant trip counts and tile sizes its loops
// unroll into straight-line code whose operand addresses are immediates (the micro-benchmark
// profiles/ubench_ring.cu reaches the 128-cycle floor per N=256 MMA exactly that way, while generic
// loops with run-time geometry cost 150-250 cycles per MMA).  tcw_matches_spec() selects it.
constexpr int kSpecKx = 96, kSpecNp = 512, kSpecOp = 80, kSpecStage = 32768, kSpecStages = 3, kSpecQA = 3, kSpecSC = 9;
inline bool tcw_matches_spec(const TcwGeom& g) {
  return g.Kx == kSpecKx && g.Np == kSpecNp && g.Op == kSpecOp && g.stage_bytes == kSpecStage && g.stages == kSpecStages &&
         g.QA == kSpecQA && g.SC == kSpecSC;
}

template <bool FP16, bool DBG, bool SPEC>
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
  const int KS_X = SPEC ? kSpecKx >> 4 : g.Kx >> 4, KH = SPEC ? kSpecNp >> 4 : g.KH, KC = SPEC ? kSpecNp >> 5 : g.Nc >> 4;
  const int NU = SPEC ? kSpecNp >> 6 : g.Nc >> 5;
  const int QA = SPEC ? kSpecQA : g.QA, SC = SPEC ? kSpecSC : g.SC, S = SPEC ? kSpecStages : g.stages;
  const int NcC = SPEC ? kSpecNp / 2 : g.Nc, OpC = SPEC ? kSpecOp : g.Op, ycolC = SPEC ? 256 - kSpecOp : g.ycol;
  const int CL = g.cluster;
  const uint16_t cl_mask = (uint16_t)((1u << CL) - 1u);
  const bool smooth = m.cost_kind == MBRL_COST_SMOOTHABS_COSH;
  const bool head = m.cost_kind == MBRL_COST_REWARD_HEAD && g.head_off >= 0;  // RewardAgent: second trunk pass + linear4 head
  const bool task = is_task_cost(m.cost_kind);
  float* const t_w4 = head ? reinterpret_cast<float*>(smem + g.head_off) : nullptr;  // [Np] fp32, zero padded
  float* const red = head ? t_w4 + g.Np : nullptr;                                    // [kTcwHeadSlots][128] per-row partial sums

  float* tab = reinterpret_cast<float*>(smem + g.tab_off);
  float *t_b3 = tab, *t_P = tab + g.Op, *t_Q = tab + 2 * g.Op, *t_sd = tab + 3 * g.Op, *t_mu = tab + 4 * g.Op;
  float *t_M = tab + 5 * g.Op;
  float *t_ainv = tab + 6 * g.Op, *t_aoff = t_ainv + kMaxAct, *costp = t_aoff + kMaxAct;
  float* xch = costp + 6 * kTcRows;  // [parity][0: a0, 1: ctl_mean][row], then picks [parity][4][row]
  uint8_t* const xa = smem + g.xa_off;
  uint8_t* const xs = smem + g.xs_off;
  uint8_t* const h2lo = smem + g.h2_off;
  const int xa_bytes = QA * 2048;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.bar_off);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + g.bar_off + 8 * kTcwBarriers);
  uint32_t* xdesc = tmem_slot + 4;  // [2 action-tile parities][kTcwMaxKx] low words of the layer-1 A descriptors
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kTcwMaxStages;
  const uint32_t bar_x = bar_empty + 8 * kTcwMaxStages, bar_acc = bar_x + 8, bar_e0 = bar_x + 16, bar_e1 = bar_x + 24;
  const uint32_t bar_l1 = bar_x + 32, bar_y = bar_x + 40;

  if (warp == kTcwMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < kTcwMaxStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, CL); }  // empty: one commit per CTA of the cluster
      mbar_init(bar_x, 1 + kTcwCostWarp0 + 4);  // sampler group + the 20 warps that write the state tile
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
  if (head) {
    for (int i = tid; i < g.Np; i += kTcwThreads) t_w4[i] = i < g.U ? __ldg(m.W4 + i) : 0.f;
    for (int i = tid; i < kTcwHeadSlots * kTcRows; i += kTcwThreads) red[i] = 0.f;  // slots of absent units stay zero
  }
  // the constant A tile of the b2 K-step: element (row, k = 0) = 1, everything else 0
  for (int i = tid; i < 256; i += kTcwThreads) {
    const uint32_t one = FP16 ? 0x3C00u : 0x3F80u;
    reinterpret_cast<uint4*>(smem + g.one_off)[i] = make_uint4(i < kTcRows ? one : 0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  if (CL > 1) cluster_sync_all();  // every CTA's barriers are initialised before any peer signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const long long R = sh.rows();

  if (warp == kTcwTmaWarp) {
    // ================= weight producer =================
    if (!(g.exp & 1) && elect_one()) {
      const uint32_t ring = smem_u32(smem + g.ring_off);
      const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
      uint32_t st = 0, ph = 0;
      for (int h = 0; h < H; ++h) {
        uint32_t off = 0;
#pragma unroll 1
        for (int pp = 0; pp < (head ? 9 : 5); ++pp) {  // reward head: W1 / W2 are streamed a second time for the reward trunk pass
          const int part = pp < 5 ? pp : pp - 5;
          if (pp == 5) off = 0;
          const uint32_t total = (uint32_t)(part < 2 ? g.p1_bytes : (part < 4 ? g.p2_bytes : g.p3_bytes));
          const uint32_t per = (uint32_t)(part < 4 ? g.tps_h * g.tile_h : g.tps_y * g.tile_y);
#pragma unroll 1
          for (uint32_t done = 0; done < total; done += per) {
            const uint32_t n = min(per, total - done);
            mbar_wait(bar_empty + 8 * st, ph ^ 1);  // first lap: passes at once on a fresh barrier
            mbar_arrive_expect_tx(bar_full + 8 * st, n);  // the whole stage: own slice + the peers' multicasts
            if (CL > 1) {
              const uint32_t slice = n / (uint32_t)CL;  // stage sizes are multiples of 512 bytes
              bulk_g2s_mc(ring + st * (uint32_t)g.stage_bytes + crank * slice, wimg + off + crank * slice, slice, bar_full + 8 * st, cl_mask);
            } else {
              bulk_g2s(ring + st * (uint32_t)g.stage_bytes, wimg + off, n, bar_full + 8 * st);
            }
            off += n;
            if (++st == (uint32_t)S) { st = 0; ph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kTcwMmaWarp) {
    // ================= MMA issuer: ONE elected lane runs the whole loop =================
    // tcgen05.mma issue blocks the issuing thread for about the MMA's own duration (queue depth ~2), so
    // whatever this thread does between two MMAs beyond ~one MMA time leaves the tensor pipe idle.
    // Measured on the first version of this loop (per stage: warp-converged wait, elect, 64-bit
    // descriptor adds, runtime-trip-count loops, __syncwarp): ~300 cycles per ring stage on top of its
    // two 128-cycle MMAs, tensor pipe 35 % busy.  Hence: descriptors as (lo, hi) words (only the
    // 14-bit start-address field of the low word changes), incremental ring bookkeeping, compile-time
    // unrolled stages on the hot layer-2 path, and the next stage's full barrier is PROBED
    // (mbarrier.test_wait) before the current stage's MMAs are issued, so its latency hides behind them.
    if (elect_one()) {
      const uint32_t idesc_h = umma_idesc(NcC, FP16), idesc_y = umma_idesc(OpC, FP16);
      const uint32_t dhi = (128u >> 4) | (1u << 14);  // SBO = 128 bytes; descriptor version 1 (bit 46)
      auto dlo = [](uint32_t saddr, uint32_t lbo_bytes) { return ((saddr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16); };
      auto d64 = [dhi](uint32_t lo) { return ((uint64_t)dhi << 32) | lo; };
      const uint32_t lo_ring_h = dlo(smem_u32(smem + g.ring_off), (uint32_t)NcC * 16);
      const uint32_t lo_ring_y = dlo(smem_u32(smem + g.ring_off), (uint32_t)OpC * 16);
      const uint32_t lo_one = dlo(smem_u32(smem + g.one_off), 2048), lo_h2 = dlo(smem_u32(h2lo), 2048);
      const uint32_t xa0 = smem_u32(xa), xs0 = smem_u32(xs);
      // The 512-column allocation necessarily starts at TMEM address 0 (checked below): using the literal
      // keeps every MMA operand address in uniform registers instead of behind the shared-memory load
      // of the allocation slot (one R2UR move per operand per MMA otherwise).
      if (tmem != 0u) __trap();
      const uint32_t tm_h = 0u, tm_acc = kTcwAccCol, tm_y = (uint32_t)ycolC;
      const uint32_t tile_h16 = SPEC ? (uint32_t)(kSpecNp / 2 * 32) >> 4 : (uint32_t)g.tile_h >> 4;
      const uint32_t tile_y16 = SPEC ? (uint32_t)(kSpecOp * 32) >> 4 : (uint32_t)g.tile_y >> 4;
      const int tps_h = SPEC ? kSpecStage / (kSpecNp / 2 * 32) : g.tps_h, tps_y = SPEC ? kSpecStage / (kSpecOp * 32) : g.tps_y;
      const uint32_t stage16 = SPEC ? (uint32_t)kSpecStage >> 4 : (uint32_t)g.stage_bytes >> 4;
      const bool use_ring = !(g.exp & 1);
      uint32_t st = 0, ph = 0, off16 = 0, ready = 0;  // ring stage, phase, offset in 16-byte units, "already full" probe
      uint32_t e0_ph = 0, e1_ph = 0;
      // layer-1 A descriptors: K-step ks of the input tile = its 8-wide chunks 2ks, 2ks+1; the action
      // chunks live in the double-buffered action tile, the state chunks in the state tile, and the
      // descriptor's LBO is simply the distance between the two chunks
      for (int par = 0; par < 2; ++par)
        for (int ks = 0; ks < KS_X; ++ks) {
          const int c0 = 2 * ks, c1 = c0 + 1;
          const uint32_t xa_cur = xa0 + (uint32_t)(par * xa_bytes);
          const uint32_t a0 = c0 < QA ? xa_cur + (uint32_t)c0 * 2048u : xs0 + (uint32_t)(c0 - QA) * 2048u;
          const uint32_t a1 = c1 < QA ? xa_cur + (uint32_t)c1 * 2048u : xs0 + (uint32_t)(c1 - QA) * 2048u;
          xdesc[par * kTcwMaxKx + ks] = dlo(a0, a1 - a0);
        }
      auto wait_epi = [&]() {  // both release rounds of the running epilogue chunk
        mbar_wait(bar_e0, e0_ph); e0_ph ^= 1;
        mbar_wait(bar_e1, e1_ph); e1_ph ^= 1;
        tc_fence_after();
      };
      auto acquire = [&]() {
        if (use_ring && (SPEC || !ready)) mbar_wait(bar_full + 8 * st, ph);
        tc_fence_after();
        const bool wrap = st + 1 == (uint32_t)S;
        // probe the next stage early (generic instantiation only: in the unrolled SPEC code the thread runs far
        // enough ahead of the tensor pipe that the probe is pure overhead: 7.43 -> 7.32 ms without it)
        ready = (SPEC || (g.exp & 8)) ? 0u : (use_ring ? mbar_test(bar_full + (wrap ? 0u : 8 * st + 8), wrap ? ph ^ 1 : ph) : 1u);
      };
      auto release = [&]() {
        if (use_ring) { if (CL > 1) tc_commit_mc(bar_empty + 8 * st, cl_mask); else tc_commit(bar_empty + 8 * st); }
        if (++st == (uint32_t)S) { st = 0; off16 = 0; ph ^= 1; } else off16 += stage16;
      };

      const int passes = head ? 2 : 1;
      for (int h = 0; h < H; ++h) {
       // pass 0: the dynamics step.  pass 1 (reward head only): RewardAgent's cost (src/mbrl/agents.py:349-358),
       // the trunk evaluated AGAIN at (s_{h+1}, a_h) -- the input tile is the action chunks of THIS step
       // (parity h) with the state chunks the output epilogue has just written -- followed by
       // ModelWithReward's linear4 head (models.py:135-141), a dot product in the layer-2 epilogue (fp32).
#pragma unroll 1
       for (int pass = 0; pass < passes; ++pass) {
        const bool dynamics = pass == 0;
        if (dynamics && head && h > 0) {
          wait_epi();  // x of this step was awaited by the reward pass of step h-1; ACC drained by its last epilogue
        } else {
          mbar_wait(bar_x, (h + pass) & 1);
          tc_fence_after();
        }
        if (DBG && dynamics) tc_stamp(dbg, h, 0);
        const uint32_t* xd = xdesc + (h & 1) * kTcwMaxKx;
        // ---- layer 1, two chunks: ACC = x . W1[chunk]^T ----
#pragma unroll (SPEC ? 16 : 1)
        for (int c = 0; c < 2; ++c) {
          if (c == 1) wait_epi();  // ACC drained by the epilogue of chunk 0
          if (DBG && dynamics) tc_stamp(dbg, h, 1 + 2 * c);
#pragma unroll (SPEC ? 16 : 1)
          for (int t0 = 0; t0 < KS_X; t0 += tps_h) {
            const int n = min(tps_h, KS_X - t0);
            acquire();
            uint32_t blo = lo_ring_h + off16;
#pragma unroll (SPEC ? 4 : 1)
            for (int i = 0; i < n; ++i, blo += tile_h16)
              mma_ss(tm_acc, d64(xd[t0 + i]), d64(blo), idesc_h, (t0 + i) > 0);
            release();
          }
          tc_commit(bar_acc);
          if (c == 1 && dynamics) tc_commit(bar_l1);  // the action tile of this step is free again
          if (DBG && dynamics) tc_stamp(dbg, h, 2 + 2 * c);
        }
        // ---- layer 2, two chunks: ACC = h1 . W2[chunk]^T + b2 (KH TS K-steps, then the b2 K-step) ----
#pragma unroll (SPEC ? 16 : 1)
        for (int c = 0; c < 2; ++c) {
          wait_epi();  // c == 0: h1 complete in TMEM; c == 1: ACC drained (dynamics: h2's lower half is in shared memory)
          if (DBG && dynamics) tc_stamp(dbg, h, 5 + 2 * c);
          int ks = 0;
          uint32_t a = tm_h;
          if (tps_h == 2) {  // two 8 KB tiles per stage, straight-line
#pragma unroll (SPEC ? 16 : 1)
            for (; ks + 2 <= KH; ks += 2, a += 16) {
              acquire();
              const uint32_t blo = lo_ring_h + off16;
              mma_ts(tm_acc, a, d64(blo), idesc_h, ks > 0);
              mma_ts(tm_acc, a + 8, d64(blo + tile_h16), idesc_h, 1);
              release();
            }
          } else if (tps_h == 4) {
#pragma unroll (SPEC ? 16 : 1)
            for (; ks + 4 <= KH; ks += 4, a += 32) {
              acquire();
              const uint32_t blo = lo_ring_h + off16;
              mma_ts(tm_acc, a, d64(blo), idesc_h, ks > 0);
              mma_ts(tm_acc, a + 8, d64(blo + tile_h16), idesc_h, 1);
              mma_ts(tm_acc, a + 16, d64(blo + 2 * tile_h16), idesc_h, 1);
              mma_ts(tm_acc, a + 24, d64(blo + 3 * tile_h16), idesc_h, 1);
              release();
            }
          }
#pragma unroll (SPEC ? 16 : 1)
          while (ks <= KH) {  // remaining stages (always the one that ends with the b2 tile)
            const int n = min(tps_h, KH + 1 - ks);
            acquire();
            uint32_t blo = lo_ring_h + off16;
#pragma unroll (SPEC ? 4 : 1)
            for (int i = 0; i < n; ++i, ++ks, blo += tile_h16) {
              if (ks < KH) { mma_ts(tm_acc, a, d64(blo), idesc_h, ks > 0); a += 8; }
              else mma_ss(tm_acc, d64(lo_one), d64(blo), idesc_h, 1);
            }
            release();
          }
          tc_commit(bar_acc);
          if (DBG && dynamics) tc_stamp(dbg, h, 6 + 2 * c);
        }
        if (!dynamics) continue;
        // ---- layer 3: y = h2 . W3^T; lower half from shared memory, upper half from TMEM as released ----
        {
          // N = Op MMAs are short (~26 + 0.43*Op cycles): the per-MMA bookkeeping is one counter
          int ks = 0, rem = KH, left = 0;
          uint32_t alo = lo_h2, a = tm_h, blo = 0;
          auto tile_begin = [&]() { if (left == 0) { acquire(); left = min(tps_y, rem); rem -= left; blo = lo_ring_y + off16; } };
          auto tile_end = [&]() { blo += tile_y16; if (--left == 0) release(); };
#pragma unroll (SPEC ? 16 : 1)
          for (; ks < KC; ++ks, alo += 4096u >> 4) { tile_begin(); mma_ss(tm_y, d64(alo), d64(blo), idesc_y, ks > 0); tile_end(); }
          // the chunk-1 epilogue releases h2's upper half in two rounds: K-steps [KC, KC+8) and the rest
          if (DBG) tc_stamp(dbg, h, 26);
          mbar_wait(bar_e0, e0_ph); e0_ph ^= 1; tc_fence_after();
          if (DBG) tc_stamp(dbg, h, 27);
          const int r0_end = min(KC + 8, KH);
#pragma unroll (SPEC ? 16 : 1)
          for (; ks < r0_end; ++ks, a += 8) { tile_begin(); mma_ts(tm_y, a, d64(blo), idesc_y, 1); tile_end(); }
          if (DBG) tc_stamp(dbg, h, 28);
          mbar_wait(bar_e1, e1_ph); e1_ph ^= 1; tc_fence_after();  // (a round without units still completes)
          if (DBG) tc_stamp(dbg, h, 29);
#pragma unroll (SPEC ? 16 : 1)
          for (; ks < KH; ++ks, a += 8) { tile_begin(); mma_ts(tm_y, a, d64(blo), idesc_y, 1); tile_end(); }
          tc_commit(bar_y);
          if (DBG) tc_stamp(dbg, h, 9);
        }
       }
      }
    }
    __syncwarp();
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
      if (DBG && srow == 0) tc_stamp(dbg, hs, 13);
      float acc = 0.f, ctl = 0.f, first = 0.f;
      float* aout = (actions_out && valid) ? actions_out + ((long long)hs * R + row) * A : nullptr;
      uint8_t* xt = xa + (hs & 1) * xa_bytes;
      const float* pm = staged ? ms_mu + ms_env + hs * A : nullptr;
      const float* ps = staged ? ms_sd + ms_env + hs * A : nullptr;
      for (int q = 0; q < QA; ++q) {
        float v[8];
        {
          float t4[4], u4[4];
          if (g.exp & 4) { t4[0] = t4[1] = t4[2] = t4[3] = 0.25f; }
          else if (8 * q < A) raw_action4(src, A, H, hs, env_l, cand_l, row, R, 2 * q, t4, pm, ps);
          else { t4[0] = t4[1] = t4[2] = t4[3] = 0.f; }
          if (g.exp & 4) { u4[0] = u4[1] = u4[2] = u4[3] = 0.25f; }
          else if (8 * q + 4 < A) raw_action4(src, A, H, hs, env_l, cand_l, row, R, 2 * q + 1, u4, pm, ps);
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
      if (srow == 0) { mbar_arrive(bar_x); if (DBG) tc_stamp(dbg, hs, 14); }
    }
    if (head && srow == 0) mbar_arrive(bar_x);  // the reward pass of the last step waits for one more x phase
    costp[5 * kTcRows + srow] = act_total;
  } else {
    // ================= hidden-epilogue warps 0-15 and cost warps 16-19 =================
    // All 20 warps share the OUTPUT epilogue: five warps own each TMEM lane quarter (warp & 3) and
    // split the y columns in 16-column groups (group index = warp >> 2, +5, ...), so the y -> cost ->
    // next-input-tile hand-over -- which sits on the step's critical path -- costs one group per
    // thread instead of all Op columns in one thread.
    const int yw = warp >> 2, quarter = warp & 3;
    const int trow = quarter * 32 + lane;
    const long long row = (long long)blockIdx.x * kTcRows + trow;
    const bool valid = row < R;
    const int env_l = valid ? (int)(row / sh.N) : 0;
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    const int NG = OpC >> 4;
    float st_total = 0.f;
    pdl_wait();
    // step 0: the normalised initial state, this warp's chunks of the state tile
    for (int gi = yw; gi < NG; gi += 5) {
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = 2 * gi + jj;
        if (j < SC) {
          float xn[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int o = 8 * j + i;
            xn[i] = (o < O && valid) ? (dep_load(s0 + (long long)env_l * O + o) - t_mu[o]) / t_sd[o] : 0.f;
          }
          uint4 pk;
          pk.x = pack2<FP16>(xn[0], xn[1]); pk.y = pack2<FP16>(xn[2], xn[3]);
          pk.z = pack2<FP16>(xn[4], xn[5]); pk.w = pack2<FP16>(xn[6], xn[7]);
          *reinterpret_cast<uint4*>(xs + j * 2048 + trow * 16) = pk;
        }
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_x);

    // One accumulator chunk (256 columns at hidden 512) through this warp's two 32-column units (release
    // rounds 0 and 1), processed as four 16-column halves through two register buffers so that a TMEM load
    // is always in flight while the previous half is converted and stored.
    //   mode 0: relu -> 16 bit -> TMEM columns dstcol.. (the next layer's A operand)
    //   mode 1: relu -> 16 bit -> h2's shared-memory A tile
    //   mode 2: reward head: per-row partial dot products with W4[wcol..] into red[slot0 + unit]
    auto hidden_chunk = [&](int h, int parity, int mode, int dstcol, int dbg_layer, int dbg_col0, int wcol, int slot0, int ev) {
      const int wg = yw;
      if (g.exp & 16) mbar_wait_nohint(bar_acc, parity); else mbar_wait(bar_acc, parity);
      tc_fence_after();
      if (DBG && tid == 0 && ev >= 0) tc_stamp(dbg, h, ev);
      const int u0 = wg, u1 = 4 + wg;
      const bool has0 = u0 < NU && !(g.exp & 2), has1 = u1 < NU && !(g.exp & 2);
      const uint32_t src0 = lane_base + (uint32_t)(kTcwAccCol + 32 * u0), src1 = lane_base + (uint32_t)(kTcwAccCol + 32 * u1);
      const uint32_t dst = lane_base + (uint32_t)dstcol;
      const float* w4 = t_w4 + wcol;
      uint32_t va[16], vb[16];
      float hacc = 0.f;
      if (has0) { tmem_ld16(src0, va); tmem_ld16(src0 + 16, vb); tmem_ld_wait(); }
      if (has0) {
        tcw_epi_half<FP16, DBG>(va, mode, dst + 16 * u0, h2lo + (4 * u0) * 2048 + trow * 16, dbg, h, trow, dbg_layer, dbg_col0 + 32 * u0, w4 + 32 * u0, hacc);
        if (has1) tmem_ld16(src1, va);
        tcw_epi_half<FP16, DBG>(vb, mode, dst + 16 * u0 + 8, h2lo + (4 * u0 + 2) * 2048 + trow * 16, dbg, h, trow, dbg_layer, dbg_col0 + 32 * u0 + 16, w4 + 32 * u0 + 16, hacc);
        if (has1) tmem_ld16(src1 + 16, vb);
        if (mode == 2) { red[(slot0 + u0) * kTcRows + trow] = hacc; hacc = 0.f; }
        else if (mode == 1) fence_proxy_async();
        else tmem_st_wait();
        tc_fence_before();
      } else if (has1) { tmem_ld16(src1, va); tmem_ld16(src1 + 16, vb); }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_e0);
      if (has1) {
        tmem_ld_wait();
        tcw_epi_half<FP16, DBG>(va, mode, dst + 16 * u1, h2lo + (4 * u1) * 2048 + trow * 16, dbg, h, trow, dbg_layer, dbg_col0 + 32 * u1, w4 + 32 * u1, hacc);
        tcw_epi_half<FP16, DBG>(vb, mode, dst + 16 * u1 + 8, h2lo + (4 * u1 + 2) * 2048 + trow * 16, dbg, h, trow, dbg_layer, dbg_col0 + 32 * u1 + 16, w4 + 32 * u1 + 16, hacc);
        if (mode == 2) red[(slot0 + u1) * kTcRows + trow] = hacc;
        else if (mode == 1) fence_proxy_async();
        else tmem_st_wait();
        tc_fence_before();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_e1);
      if (DBG && tid == 0 && ev >= 0) tc_stamp(dbg, h, ev + 1);
    };
    float head_total = 0.f;  // reward-head cost accumulated by the cost warps

    for (int h = 0; h < H; ++h) {
     // pass 0: the dynamics step (4 accumulator chunks, then the output epilogue); pass 1 (reward head only):
     // the reward trunk pass (4 more chunks: L1' packs h1', L2' feeds the linear4 dot product)
#pragma unroll 1
     for (int pass = 0; pass < (head ? 2 : 1); ++pass) {
      if (warp < kTcwEpiWarps) {
        // accumulator completions come in fours (L1 chunk 0, L1 chunk 1, L2 chunk 0, L2 chunk 1): parity q & 1
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          // mode: L1 chunks and the dynamics L2 chunk 1 pack into TMEM, the dynamics L2 chunk 0 into h2's
          // shared-memory tile, the reward pass's L2 chunks feed the head
          const int mode = q < 2 ? 0 : (pass == 1 ? 2 : (q == 2 ? 1 : 0));
          const int dstcol = q == 1 ? NcC >> 1 : 0;
          // the cost warp of this lane quarter has consumed the head partial sums of step h-1
          if (pass == 1 && q == 2 && h > 0) asm volatile("bar.sync %0, 160;" ::"r"(7 + quarter) : "memory");
          hidden_chunk(h, q & 1, mode, dstcol, pass == 0 ? (q >> 1) : -1, (q & 1) * NcC, (q & 1) * NcC, (q & 1) * (kTcwHeadSlots / 2),
                       pass == 0 ? 16 + 2 * q : -1);
        }
      }
      if (pass == 1) {
        // ---- RewardAgent's cost: linear4 head of the second trunk pass (agents.py:349-358, models.py:135-141) ----
        // Producer / consumer named barriers per lane quarter (4 epilogue warps + 1 cost warp = 160 threads): the
        // side that only SIGNALS uses bar.arrive -- a cost warp blocked in bar.sync on the "consumed" barrier
        // could not take part in the next step's output epilogue, and the step would never get its x-ready.
        if (yw == 4) {
          asm volatile("bar.sync %0, 160;" ::"r"(3 + quarter) : "memory");  // the quarter's partial sums of step h are written
          float r = 0.f;
#pragma unroll
          for (int sl = 0; sl < kTcwHeadSlots; ++sl) r += red[sl * kTcRows + trow];
          head_total += fmaf(r + m.b4, m.sd_r, m.mu_r);  // unnormalize_field(linear4(h2)) (data.py:255-257); minimised as the reference does
          if (h + 1 < H) asm volatile("bar.arrive %0, 160;" ::"r"(7 + quarter) : "memory");  // consumed
        } else {
          asm volatile("bar.arrive %0, 160;" ::"r"(3 + quarter) : "memory");
        }
        continue;
      }
      // ---- output epilogue: y + b3 (fp32), un-normalise, cost, next input tile ----
      if (g.exp & 16) mbar_wait_nohint(bar_y, h & 1); else mbar_wait(bar_y, h & 1);
      tc_fence_after();
      if (DBG && tid == 0) tc_stamp(dbg, h, 11);
      float* pk_slot = xch + 4 * kTcRows + (h & 1) * 4 * kTcRows;  // task-cost picks [parity][4][row]
      // Critical path first: y -> next step's state tile -> x-ready; the cost of step h is accumulated
      // afterwards, under the next step's layer-1 MMAs.  (Rows of W3 beyond O are zero and b3 is
      // zero-padded, so padded y columns are exactly 0 without a mask.)
      const bool one_group = NG <= 5;  // every warp owns at most one 16-column group: keep its y in registers
      uint32_t v[16];
      if (one_group) {
        if (yw < NG) {
          tmem_ld16(lane_base + (uint32_t)(ycolC + 16 * yw), v);
          tmem_ld_wait();
        }
      }
      if (h + 1 < H || head) {  // (the reward pass of the last step still needs the state tile)
#pragma unroll 1
        for (int gi = yw; gi < NG; gi += 5) {
          if (!one_group) { tmem_ld16(lane_base + (uint32_t)(ycolC + 16 * gi), v); tmem_ld_wait(); }
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int j = 2 * gi + jj;
            if (j < SC) {
              const float4 ba = *reinterpret_cast<const float4*>(t_b3 + 16 * gi + 8 * jj), bb = *reinterpret_cast<const float4*>(t_b3 + 16 * gi + 8 * jj + 4);
              uint4 pk;
              pk.x = pack2<FP16>(__uint_as_float(v[8 * jj + 0]) + ba.x, __uint_as_float(v[8 * jj + 1]) + ba.y);
              pk.y = pack2<FP16>(__uint_as_float(v[8 * jj + 2]) + ba.z, __uint_as_float(v[8 * jj + 3]) + ba.w);
              pk.z = pack2<FP16>(__uint_as_float(v[8 * jj + 4]) + bb.x, __uint_as_float(v[8 * jj + 5]) + bb.y);
              pk.w = pack2<FP16>(__uint_as_float(v[8 * jj + 6]) + bb.z, __uint_as_float(v[8 * jj + 7]) + bb.w);
              *reinterpret_cast<uint4*>(xs + j * 2048 + trow * 16) = pk;
            }
          }
        }
        fence_proxy_async();   // generic-proxy writes of the state tile -> visible to the MMA
        if (one_group) {
          tc_fence_before();   // our tcgen05.ld of y is ordered before the columns are reused
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_x);
          if (DBG && tid == 0) tc_stamp(dbg, h, 12);
        }
      }
#pragma unroll 1
      for (int gi = yw; gi < NG; gi += 5) {
        if (!one_group) { tmem_ld16(lane_base + (uint32_t)(ycolC + 16 * gi), v); tmem_ld_wait(); }
        if (DBG && dbg && blockIdx.x == 0 && h == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) dbg[(2 * kTcRows + trow) * kTcDbgCols + 16 * gi + i] = __uint_as_float(v[i]);
        }
        if (smooth) {
          float term[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int o = 16 * gi + i;  // < Op: tables are Op long, padded entries are zero / masked
            const float x = fmaf(__uint_as_float(v[i]), t_P[o], t_Q[o]);  // (s - goal) * w with s = (y_raw + b3)*sd + mu
            term[i] = (fast_sqrt(fmaf(x, x, m.alpha2)) - m.alpha) * t_M[o];
          }
          st_total += ((term[0] + term[1]) + (term[2] + term[3])) + ((term[4] + term[5]) + (term[6] + term[7])) +
                      (((term[8] + term[9]) + (term[10] + term[11])) + ((term[12] + term[13]) + (term[14] + term[15])));
        } else if (task) {
          int pick[4];
          task_pick_indices(m.cost_kind, pick);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int o = 16 * gi + i;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (o == pick[q]) pk_slot[q * kTcRows + trow] = fmaf(__uint_as_float(v[i]) + t_b3[o], t_sd[o], t_mu[o]);
          }
        }
        if (states_out && valid) {
          float* sout = states_out + ((long long)h * R + row) * O;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int o = 16 * gi + i;
            if (o < O) sout[o] = fmaf(__uint_as_float(v[i]) + t_b3[o], t_sd[o], t_mu[o]);  // unnormalize_state (data.py:255-257)
          }
        }
      }
      if ((h + 1 < H || head) && !one_group) {
        tc_fence_before();     // our tcgen05.ld of y is ordered before the columns are reused
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_x);
        if (DBG && tid == 0) tc_stamp(dbg, h, 12);
      }
      if (task) {
        // the five warps of this lane quarter have stored the picked state entries of step h
        asm volatile("bar.sync %0, 160;" ::"r"(3 + quarter) : "memory");
        if (yw == 4) {
          int pick[4];
          task_pick_indices(m.cost_kind, pick);
          float p4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) p4[q] = pick[q] >= 0 ? pk_slot[q * kTcRows + trow] : 0.f;
          st_total += task_cost(m.cost_kind, p4, xch[((h & 1) * 2 + 0) * kTcRows + trow], xch[((h & 1) * 2 + 1) * kTcRows + trow]);
        }
      }
     }
    }
    costp[yw * kTcRows + trow] = valid ? st_total + head_total : 0.f;
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all();  // no CTA exits while a peer may still signal its ring barriers
  else __syncthreads();
  tc_fence_after();
  if (tid < kTcRows) {
    const long long row = (long long)blockIdx.x * kTcRows + tid;
    if (row < R) costs[row] = (((costp[tid] + costp[kTcRows + tid]) + (costp[2 * kTcRows + tid] + costp[3 * kTcRows + tid])) + costp[4 * kTcRows + tid]) + costp[5 * kTcRows + tid];
  }
  if (warp == kTcwMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace mbrl
