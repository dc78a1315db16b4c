// tcgen05 tensor-core rollout + cost engine (MBRL_ENGINE_TC_FP16 / MBRL_ENGINE_TC_BF16).
//
// Same fused computation as rollout_simt.cuh -- the hot loop of
// RandomShootingPlanner._generate_trajectories (src/mbrl/planners.py:199-210) with
// DynamicsModel.forward (src/mbrl/models.py:13-29), the 3-layer ReLU MLP (models.py:106-110),
// the normalisers (src/mbrl/data.py:255-260) and SmoothAbs+Cosh cost (models.py:244-272) --
// but the three layer GEMMs of every step run on the 5th-generation tensor cores:
//
//   * one CTA owns a tile of 128 candidate rows for all H steps (row r <-> TMEM lane r);
//   * the three weight matrices are packed once on the host into the UMMA canonical K-major
//     (no-swizzle, 8x16B core matrices) layout as 16-bit operands and TMA-bulk-copied
//     (cp.async.bulk) into shared memory at kernel start, where they stay for the whole
//     rollout;
//   * layer 1: SS-mode tcgen05.mma, A = the 128 x Kx input tile in shared memory
//     ([action section | 1 | state section], written by the epilogue threads),
//     D1 -> TMEM columns [0, Np);
//   * epilogue 1: tcgen05.ld D1 chunk -> cvt.rn.relu.{f16,bf16}x2 -> tcgen05.st the packed
//     activations back into the first half of the chunk's OWN columns (no cross-warp overlap);
//     each finished 32-column chunk releases
//     the next layer's K-steps through an mbarrier, so the layer-2 MMAs (TS mode: A straight
//     from TMEM, D2 -> columns [256, 256+Np)) overlap the rest of the epilogue;
//   * epilogue 2 / layer 3 the same way (h2 in place at [256, 256+Np/2), D3 -> [0, Op));
//   * epilogue 3: D3 + b3 in fp32 -> un-normalise -> accumulate the cost -> the normalised
//     prediction is the next step's input tile.  b1/b2 ride in the GEMMs through a constant-1
//     input column (K padding that exists anyway), so the hidden epilogues are pure
//     load/convert/store.
//   * actions come from the Philox sampler (or injected buffers) one step ahead, hidden
//     behind the layer-1 MMA.
//
// Nothing but costs[R] leaves the SM (debug trajectory outputs are optional).
// Operand precision: fp16 or bf16 operands, fp32 accumulation (TMEM), fp32 bias-3,
// un-normalisation and cost.  Tolerances are stated in tests/test_gpu_tc.py.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "philox.cuh"

namespace mbrl {

constexpr int kTcRows = 128;           // MMA M
constexpr int kTcEpiWarps = 8;         // warps 0-7: two warpgroups share the 128 TMEM lanes
constexpr int kTcSampWarps = 4;        // warps 8-11: one sampler thread per row
constexpr int kTcMmaWarp = 12;         // warp 12: TMEM alloc, weight TMA, MMA issue
constexpr int kTcThreads = 13 * 32;
constexpr int kTcD2Col = 256;          // TMEM column of the layer-2 accumulator
constexpr int kTcMaxChunks = 8;        // ceil(256 / 32)

struct TcGeom {
  int O, A, U;
  int Ka;  // action section of the input tile: A actions, the constant 1, zero pad (multiple of 8)
  int Kx;  // layer-1 K (multiple of 16): Ka + O rounded up
  int Np;  // padded hidden width (multiple of 16, > U: column U carries the constant 1)
  int Op;  // layer-3 N (multiple of 32)
  int w1_off, w2_off, w3_off, w_bytes;  // packed operand image
  int x_bytes;                           // one input tile
  int tab_off, x_off, bar_off, smem_bytes;
};

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

inline bool tc_geometry(int O, int A, int U, size_t max_smem, TcGeom* g, std::string* why) {
  g->O = O; g->A = A; g->U = U;
  g->Ka = round_up(A + 1, 8);
  g->Kx = round_up(g->Ka + O, 16);
  g->Np = round_up(U + 1, 16);
  g->Op = round_up(O, 32);
  if (g->Np > 256) { *why = "hidden > 255 needs the N-split / weight-streaming kernel (not built yet)"; return false; }
  if (g->Op > 128) { *why = "obs_dim > 128 unsupported"; return false; }
  g->w1_off = 0;
  g->w2_off = g->w1_off + g->Kx * g->Np * 2;
  g->w3_off = g->w2_off + g->Np * g->Np * 2;
  g->w_bytes = g->w3_off + g->Np * g->Op * 2;
  g->x_bytes = g->Kx * kTcRows * 2;
  g->tab_off = g->w_bytes;  // fp32: 5 tables of Op, 2 of kMaxAct, 3 x 128 cost partials
  g->x_off = round_up(g->tab_off + (5 * g->Op + 2 * kMaxAct + 3 * kTcRows) * 4, 128);
  g->bar_off = g->x_off + 2 * g->x_bytes;
  g->smem_bytes = g->bar_off + 8 * (5 + 2 * kTcMaxChunks) + 16;
  if ((size_t)g->smem_bytes > max_smem) { *why = "weights do not fit shared memory (streaming kernel not built yet)"; return false; }
  return true;
}

// ---- host-side packing -----------------------------------------------------------------
// Canonical K-major no-swizzle operand: element (row n, k) at byte
//   (k/8) * rows*16 + n*16 + (k%8)*2          (8x16B core matrices, LBO = rows*16, SBO = 128)
inline void tc_put(std::vector<uint16_t>& img, int off_bytes, int rows, int n, int k, float v, bool fp16) {
  uint16_t bits;
  if (fp16) { __half h = __float2half_rn(v); std::memcpy(&bits, &h, 2); }
  else { __nv_bfloat16 h = __float2bfloat16_rn(v); std::memcpy(&bits, &h, 2); }
  img[(size_t)off_bytes / 2 + (size_t)(k / 8) * rows * 8 + (size_t)n * 8 + (k % 8)] = bits;
}

// W1 [U, O+A], W2 [U, U], W3 [O, U] in nn.Linear layout (src/mbrl/models.py:99-101).
inline void tc_pack(const TcGeom& g, bool fp16, const float* W1, const float* b1, const float* W2, const float* b2,
                    const float* W3, std::vector<uint16_t>* out) {
  const int O = g.O, A = g.A, U = g.U, D = O + A;
  std::vector<uint16_t>& img = *out;
  img.assign((size_t)g.w_bytes / 2, 0);
  // layer 1: input column e:  e < A -> action e;  e == A -> constant 1;  Ka <= e < Ka+O -> state e-Ka
  for (int n = 0; n < U; ++n) {
    for (int a = 0; a < A; ++a) tc_put(img, g.w1_off, g.Np, n, a, W1[(size_t)n * D + O + a], fp16);
    tc_put(img, g.w1_off, g.Np, n, A, b1[n], fp16);
    for (int o = 0; o < O; ++o) tc_put(img, g.w1_off, g.Np, n, g.Ka + o, W1[(size_t)n * D + o], fp16);
  }
  tc_put(img, g.w1_off, g.Np, U, A, 1.0f, fp16);  // hidden unit U == relu(1) == 1 carries b2
  for (int n = 0; n < U; ++n) {
    for (int k = 0; k < U; ++k) tc_put(img, g.w2_off, g.Np, n, k, W2[(size_t)n * U + k], fp16);
    tc_put(img, g.w2_off, g.Np, n, U, b2[n], fp16);
  }
  for (int o = 0; o < O; ++o)
    for (int k = 0; k < U; ++k) tc_put(img, g.w3_off, g.Op, o, k, W3[(size_t)o * U + k], fp16);
}

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // try_wait with a suspend-time hint: the waiting thread sleeps in hardware until the phase
  // completes (or the hint expires) instead of spinning -- with ~20 mostly-waiting warps per CTA,
  // spinning waiters would take issue slots from the warp that issues the MMAs.
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(bar), "r"(parity), "r"(0x989680) : "memory");
}
// try_wait WITHOUT a suspend-time hint: the instruction itself blocks in hardware until the phase completes
// or an implementation-defined time limit passes (the CUTLASS ClusterBarrier::wait loop).
__device__ __forceinline__ void mbar_wait_nohint(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// One elected lane of a fully converged warp (what cute::elect_one_sync emits).  Code guarded by
// this predicate is known to the compiler to run in exactly one thread, so warp-level
// instructions (UTCHMMA, UTCBAR, UBLKCP) are emitted once with uniform-register operands instead
// of inside a per-active-thread loop with R2UR moves.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// K-major, no swizzle, descriptor version 1 (Blackwell).  Fields are in 16-byte units.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
         ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: fp32 accumulate, A/B K-major, M=128.
__host__ __device__ inline uint32_t umma_idesc(int n, bool fp16) {
  const uint32_t fmt = fp16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

#define MBRL_R8(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
#define MBRL_I8(v, o) "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7])

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : MBRL_R8(v, 0), MBRL_R8(v, 8), MBRL_R8(v, 16), MBRL_R8(v, 24)
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : MBRL_R8(v, 0), MBRL_R8(v, 8)
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : MBRL_R8(v, 0), MBRL_R8(v, 8)
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      MBRL_I8(v, 0), MBRL_I8(v, 8) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), MBRL_I8(v, 0) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// relu + round-to-nearest + pack: lo -> bits [0,16) (even k), hi -> bits [16,32) (odd k)
template <bool FP16>
__device__ __forceinline__ uint32_t pack_relu(uint32_t lo_bits, uint32_t hi_bits) {
  uint32_t d;
  if (FP16)
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi_bits)), "f"(__uint_as_float(lo_bits)));
  else
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi_bits)), "f"(__uint_as_float(lo_bits)));
  return d;
}
template <bool FP16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t d;
  if (FP16)
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// ---- action sampling, 4 raw actions for dims 4g..4g+3 (0 beyond A) --------------------------
// ms_mu / ms_sd: optional shared-memory copy of this tile's mean/std rows, already offset to
// (env, h) (see the fused kernel's sampler); null -> coherent global loads (dep_load: the mean/std
// were written by the refit kernel that precedes this launch).
// raw_noise4: the draw itself (depends only on counters / the static injected buffer -- may run
// before griddepcontrol.wait); apply_action4: mean/std/clip (reads the predecessor's refit).
__device__ __forceinline__ void raw_noise4(const ActionSource& s, int A, int H, int h, int env_l, int cand_l,
                                           long long row, long long R, int g, float (&z)[4]) {
  if (s.mode == MBRL_SAMPLE_INJECT_ACTIONS || s.mode == MBRL_SAMPLE_INJECT_NOISE) {
    const float* p = s.buf + ((long long)h * R + row) * A;
#pragma unroll
    for (int j = 0; j < 4; ++j) z[j] = __ldg(p + min(4 * g + j, A - 1));
  } else {
    const int G = (A + 3) >> 2;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)(h * G + g), s.iteration, s.cand_offset + (uint32_t)cand_l,
                                             s.env_offset + (uint32_t)env_l),
                                  make_uint2(s.seed_lo, s.seed_hi));
    if (s.mode == MBRL_SAMPLE_GAUSSIAN) {
      const float4 q = box_muller4(r);
      z[0] = q.x; z[1] = q.y; z[2] = q.z; z[3] = q.w;
    } else {
      z[0] = u32_to_uniform(r.x); z[1] = u32_to_uniform(r.y); z[2] = u32_to_uniform(r.z); z[3] = u32_to_uniform(r.w);
    }
  }
}
__device__ __forceinline__ void apply_action4(const ActionSource& s, int A, int H, int h, int env_l, int g,
                                              const float (&z)[4], float (&out)[4],
                                              const float* ms_mu = nullptr, const float* ms_sd = nullptr) {
  // branch-free per element: indices are clamped into range and the result is masked, so the four
  // independent load -> fma -> clip chains can overlap
  const long long ms = ((long long)env_l * H + h) * A;
  const bool affine = s.mode == MBRL_SAMPLE_INJECT_NOISE || s.mode == MBRL_SAMPLE_GAUSSIAN;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int a = 4 * g + j, ac = min(a, A - 1);
    float v = z[j];
    if (affine) {
      const float mu = ms_mu ? ms_mu[ac] : dep_load(s.mu + ms + ac);
      const float sd = ms_sd ? ms_sd[ac] : dep_load(s.sd + ms + ac);
      v = clipf(__fadd_rn(mu, __fmul_rn(sd, v)), s.lo, s.hi);
    }
    else if (s.mode == MBRL_SAMPLE_UNIFORM) v = __fadd_rn(s.lo, __fmul_rn(__fsub_rn(s.hi, s.lo), v));
    out[j] = a < A ? v : 0.f;
  }
}
__device__ __forceinline__ void raw_action4(const ActionSource& s, int A, int H, int h, int env_l, int cand_l,
                                            long long row, long long R, int g, float (&out)[4],
                                            const float* ms_mu = nullptr, const float* ms_sd = nullptr) {
  float z[4];
  raw_noise4(s, A, H, h, env_l, cand_l, row, R, g, z);
  apply_action4(s, A, H, h, env_l, g, z, out, ms_mu, ms_sd);
}

// ---- the kernel ----------------------------------------------------------------------------
// Warp roles (416 threads, 1 CTA per SM):
//   warps 0-7   epilogue: thread <-> TMEM lane (warp%4)*32+lane; the two warpgroups split the
//               accumulator columns (even / odd 32-column chunks; 16-column halves of D3)
//   warps 8-11  sampler: one thread per row draws the next step's actions (Philox + Box-Muller
//               + clip), normalises them into the input tile and accumulates the action cost --
//               entirely off the GEMM critical path
//   warp 12     one elected thread: weight TMA, all tcgen05.mma issue, commits
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float cosh_m1_fast(float t) {
  const float e = __expf(t);
  return 0.5f * (e + __fdividef(1.0f, e)) - 1.0f;
}

// Diagnostic timeline (tests/profiling only): when the debug buffer is armed, tile 1 records
// clock64() at the hand-over points of every step: slot [h][e], e = 0..15.
constexpr int kTcDbgCols = 512;  // dump row pitch (the wide engine has 512 hidden columns)
constexpr int kTcDbgFloats = 3 * kTcRows * kTcDbgCols;
constexpr int kTcTimelineSteps = 64, kTcTimelineEvents = 32;
__device__ __forceinline__ void tc_stamp(float* dbg, int h, int e) {
  if (dbg && blockIdx.x == 1 && h < kTcTimelineSteps)
    reinterpret_cast<long long*>(dbg + kTcDbgFloats)[h * kTcTimelineEvents + e] = clock64();
}

template <bool FP16>
__global__ void __launch_bounds__(kTcThreads, 1)
rollout_tc_kernel(TcGeom g, const uint8_t* __restrict__ wimg, ModelDev m, ActionSource src, Shape sh,
                  const float* __restrict__ s0, float* __restrict__ costs, float* __restrict__ states_out,
                  float* __restrict__ actions_out, float* __restrict__ dbg) {
  extern __shared__ __align__(128) uint8_t tc_smem[];
  uint8_t* const smem = tc_smem;
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int O = g.O, A = g.A, H = sh.H;
  const int NC = (g.Np + 31) >> 5;       // hidden epilogue chunks (32 columns, last may be 16)
  const int KS_H = g.Np >> 4;            // K-steps of layers 2 and 3
  const int KS_X = g.Kx >> 4;            // K-steps of layer 1
  const int QA = g.Ka >> 3;              // action chunks of the input tile
  const int SC = (g.Kx - g.Ka) >> 3;     // state chunks of the input tile

  // fp32 tables: per output o: b3, P = sd*w, Q = (b3*sd + mu - goal)*w, sd, mu; per action a:
  // 1/sd_a, mu_a/sd_a; then the three per-row cost partials
  float* tab = reinterpret_cast<float*>(smem + g.tab_off);
  float* t_b3 = tab, *t_P = tab + g.Op, *t_Q = tab + 2 * g.Op, *t_sd = tab + 3 * g.Op, *t_mu = tab + 4 * g.Op;
  float* t_ainv = tab + 5 * g.Op, *t_aoff = t_ainv + kMaxAct, *costp = t_aoff + kMaxAct;
  uint8_t* xbuf = smem + g.x_off;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.bar_off);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + g.bar_off + 8 * (5 + 2 * kTcMaxChunks));
  const uint32_t bar0 = smem_u32(bars);
  // barrier map: 0 weights, 1 x-ready, 2 d1, 3 d2, 4 d3, 5.. a1[c], 5+8.. a2[c]
  const uint32_t bar_w = bar0, bar_x = bar0 + 8, bar_d1 = bar0 + 16, bar_d2 = bar0 + 24, bar_d3 = bar0 + 32;
  const uint32_t bar_a1 = bar0 + 40, bar_a2 = bar0 + 40 + 8 * kTcMaxChunks;

  if (warp == kTcMmaWarp) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
      mbar_init(bar_x, kTcEpiWarps + 1);  // 8 epilogue warps (state section) + the sampler group
      mbar_init(bar_d1, 1); mbar_init(bar_d2, 1); mbar_init(bar_d3, 1);
      for (int c = 0; c < kTcMaxChunks; ++c) { mbar_init(bar_a1 + 8 * c, 4); mbar_init(bar_a2 + 8 * c, 4); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < g.Op; i += kTcThreads) {
    const bool in = i < O;
    const float b3 = in ? __ldg(m.b3 + i) : 0.f, sd = in ? __ldg(m.sd_s + i) : 1.f, mu = in ? __ldg(m.mu_s + i) : 0.f;
    const float w = in ? __ldg(m.cost_w + i) : 0.f, goal = in ? __ldg(m.goal + i) : 0.f;
    t_b3[i] = b3; t_sd[i] = sd; t_mu[i] = mu;
    t_P[i] = sd * w;
    t_Q[i] = (b3 * sd + mu - goal) * w;
  }
  for (int i = tid; i < kMaxAct; i += kTcThreads) {
    const float inv = i < A ? 1.0f / __ldg(m.sd_a + i) : 0.f;
    t_ainv[i] = inv;
    t_aoff[i] = i < A ? __ldg(m.mu_a + i) * inv : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const long long R = sh.rows();

  if (warp == kTcMmaWarp) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_w, (uint32_t)g.w_bytes);
      bulk_g2s(smem_u32(smem), wimg, (uint32_t)g.w_bytes, bar_w);
      const uint32_t idesc_h = umma_idesc(g.Np, FP16), idesc_o = umma_idesc(g.Op, FP16);
      const uint32_t lbo_h = (uint32_t)g.Np * 16, lbo_o = (uint32_t)g.Op * 16, lbo_x = kTcRows * 16;
      // Descriptors are loop invariant; a K-step advances the 14-bit start-address field by
      // 2*LBO/16 (smem addresses stay below 256 KB, so the add never carries out of the field).
      const uint64_t d_w1 = umma_desc(smem_u32(smem + g.w1_off), lbo_h, 128);
      const uint64_t d_w2 = umma_desc(smem_u32(smem + g.w2_off), lbo_h, 128);
      const uint64_t d_w3 = umma_desc(smem_u32(smem + g.w3_off), lbo_o, 128);
      const uint64_t d_x0 = umma_desc(smem_u32(xbuf), lbo_x, 128);
      const uint64_t d_x1 = umma_desc(smem_u32(xbuf + g.x_bytes), lbo_x, 128);
      const uint64_t step_h = (2 * lbo_h) >> 4, step_o = (2 * lbo_o) >> 4, step_x = (2 * lbo_x) >> 4;
      mbar_wait(bar_w, 0);
      for (int h = 0; h < H; ++h) {
        const uint32_t ph = h & 1;
        mbar_wait(bar_x, ph);
        tc_fence_after();
        tc_stamp(dbg, h, 0);
        {
          uint64_t ad = (h & 1) ? d_x1 : d_x0, bd = d_w1;
          mma_ss(tmem, ad, bd, idesc_h, 0);
          tc_stamp(dbg, h, 20);
          for (int ks = 1; ks < KS_X; ++ks) { ad += step_x; bd += step_h; mma_ss(tmem, ad, bd, idesc_h, 1); }
        }
        tc_stamp(dbg, h, 21);
        tc_commit(bar_d1);
        tc_stamp(dbg, h, 1);
        {
          uint64_t bd = d_w2;
          uint32_t a = tmem, acc = 0;
          int left = KS_H;
          for (int c = 0; c < NC; ++c) {
            mbar_wait(bar_a1 + 8 * c, ph);
            tc_fence_after();
            mma_ts(tmem + kTcD2Col, a, bd, idesc_h, acc);
            acc = 1; bd += step_h;
            if (left > 1) { mma_ts(tmem + kTcD2Col, a + 8, bd, idesc_h, 1); bd += step_h; }
            a += 32; left -= 2;  // packed chunk c sits in the first half of its own 32 columns
          }
        }
        tc_commit(bar_d2);
        tc_stamp(dbg, h, 2);
        {
          // layer 3 (N = Op) is issue-bound, not tensor-bound: chunk-wise release hides the
          // issue cost behind the rest of epilogue 2
          uint64_t bd = d_w3;
          uint32_t a = tmem + kTcD2Col, acc = 0;
          int left = KS_H;
          for (int c = 0; c < NC; ++c) {
            mbar_wait(bar_a2 + 8 * c, ph);
            tc_fence_after();
            mma_ts(tmem, a, bd, idesc_o, acc);
            acc = 1; bd += step_o;
            if (left > 1) { mma_ts(tmem, a + 8, bd, idesc_o, 1); bd += step_o; }
            a += 32; left -= 2;
          }
        }
        tc_commit(bar_d3);
        tc_stamp(dbg, h, 3);
      }
    }
    __syncwarp();
  } else if (warp >= kTcEpiWarps) {
    // ================= sampler threads (one per row) =================
    const int srow = tid - kTcEpiWarps * 32;
    const long long row = (long long)blockIdx.x * kTcRows + srow;
    const bool valid = row < R;
    const int env_l = valid ? (int)(row / sh.N) : 0;
    const int cand_l = valid ? (int)(row - (long long)env_l * sh.N) : 0;
    const float inv_beta = 1.0f / m.beta, cscale = m.beta2 / (float)A;
    float act_total = 0.f;
    for (int hs = 0; hs < H; ++hs) {
      // the layer-1 MMA of step hs-1 has finished reading the tiles: buffer hs&1 is free and
      // the x-ready barrier has moved on to phase hs
      if (hs >= 1) mbar_wait(bar_d1, (hs - 1) & 1);
      if (srow == 0) tc_stamp(dbg, hs, 12);
      float acc = 0.f;
      float* aout = (actions_out && valid) ? actions_out + ((long long)hs * R + row) * A : nullptr;
      uint8_t* xt = xbuf + (hs & 1) * g.x_bytes;
      for (int q = 0; q < QA; ++q) {
        float v[8];
        {
          float t4[4];
          raw_action4(src, A, H, hs, env_l, cand_l, row, R, 2 * q, t4);
          v[0] = t4[0]; v[1] = t4[1]; v[2] = t4[2]; v[3] = t4[3];
          if (8 * q + 4 < A) raw_action4(src, A, H, hs, env_l, cand_l, row, R, 2 * q + 1, t4);
          else { t4[0] = t4[1] = t4[2] = t4[3] = 0.f; }
          v[4] = t4[0]; v[5] = t4[1]; v[6] = t4[2]; v[7] = t4[3];
        }
        float xn[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int a = 8 * q + i;
          if (a < A && valid) {
            acc += cosh_m1_fast(v[i] * inv_beta);
            xn[i] = fmaf(v[i], t_ainv[a], -t_aoff[a]);
            if (aout) aout[a] = v[i];
          } else {
            xn[i] = (a == A) ? 1.0f : 0.0f;
          }
        }
        uint4 pk;
        pk.x = pack2<FP16>(xn[0], xn[1]); pk.y = pack2<FP16>(xn[2], xn[3]);
        pk.z = pack2<FP16>(xn[4], xn[5]); pk.w = pack2<FP16>(xn[6], xn[7]);
        *reinterpret_cast<uint4*>(xt + q * (kTcRows * 16) + srow * 16) = pk;
      }
      act_total = fmaf(cscale, acc, act_total);  // CoshLoss: beta^2 * mean_a(cosh(a/beta) - 1)
      fence_proxy_async();
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four sampler warps
      if (srow == 0) { mbar_arrive(bar_x); tc_stamp(dbg, hs, 13); }
    }
    costp[2 * kTcRows + srow] = act_total;
  } else {
    // ================= epilogue threads =================
    const int wg = warp >> 2, quarter = warp & 3;
    const int trow = quarter * 32 + lane;                 // row in the tile == TMEM lane
    const long long row = (long long)blockIdx.x * kTcRows + trow;
    const bool valid = row < R;
    const int env_l = valid ? (int)(row / sh.N) : 0;
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    float st_total = 0.f;

    // step 0: normalised s0 into this warpgroup's state chunks
    for (int j = 0; j < SC; ++j) {
      if (((j >> 1) & 1) != wg) continue;
      float xn[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int o = 8 * j + i;
        xn[i] = (o < O && valid) ? (dep_load(s0 + (long long)env_l * O + o) - t_mu[o]) / t_sd[o] : 0.f;
      }
      uint4 pk;
      pk.x = pack2<FP16>(xn[0], xn[1]); pk.y = pack2<FP16>(xn[2], xn[3]);
      pk.z = pack2<FP16>(xn[4], xn[5]); pk.w = pack2<FP16>(xn[6], xn[7]);
      *reinterpret_cast<uint4*>(xbuf + (QA + j) * (kTcRows * 16) + trow * 16) = pk;
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_x);

    for (int h = 0; h < H; ++h) {
      const uint32_t ph = h & 1;
      // hidden epilogues: TMEM fp32 -> relu -> 16-bit, in place; chunk-wise release
#pragma unroll 1
      for (int layer = 0; layer < 2; ++layer) {
        const uint32_t dcol = layer == 0 ? 0u : (uint32_t)kTcD2Col;
        mbar_wait(layer == 0 ? bar_d1 : bar_d2, ph);
        tc_fence_after();
        if (tid == 0) tc_stamp(dbg, h, 4 + 2 * layer);
#pragma unroll 1
        for (int c = wg; c < NC; c += 4) {
          // this warpgroup's chunks c and c+2: both TMEM loads in flight before converting
          const int c2 = c + 2;
          const bool has2 = c2 < NC;
          const bool full = 32 * c + 32 <= g.Np, full2 = has2 && 32 * c2 + 32 <= g.Np;
          uint32_t v[32], v2[32], pk[16];
          if (full) tmem_ld32(lane_base + dcol + 32 * c, v);
          else tmem_ld16(lane_base + dcol + 32 * c, v);
          if (has2) {
            if (full2) tmem_ld32(lane_base + dcol + 32 * c2, v2);
            else tmem_ld16(lane_base + dcol + 32 * c2, v2);
          }
          tmem_ld_wait();
          if (tid == 0 && layer == 0 && c == 0) tc_stamp(dbg, h, 16);
          if (dbg && blockIdx.x == 0 && h == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (full || i < 16) dbg[(layer * kTcRows + trow) * kTcDbgCols + 32 * c + i] = __uint_as_float(v[i]);
              if (has2 && (full2 || i < 16)) dbg[(layer * kTcRows + trow) * kTcDbgCols + 32 * c2 + i] = __uint_as_float(v2[i]);
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_relu<FP16>(v[2 * i], v[2 * i + 1]);
          if (full) tmem_st16(lane_base + dcol + 32 * c, pk);
          else tmem_st8(lane_base + dcol + 32 * c, pk);
          if (has2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack_relu<FP16>(v2[2 * i], v2[2 * i + 1]);
            if (full2) tmem_st16(lane_base + dcol + 32 * c2, pk);
            else tmem_st8(lane_base + dcol + 32 * c2, pk);
          }
          if (tid == 0 && layer == 0 && c == 0) tc_stamp(dbg, h, 17);
          tmem_st_wait();
          tc_fence_before();
          if (tid == 0 && layer == 0 && c == 0) tc_stamp(dbg, h, 18);
          __syncwarp();
          if (lane == 0) {
            const uint32_t bar_a = layer == 0 ? bar_a1 : bar_a2;
            mbar_arrive(bar_a + 8 * c);
            if (has2) mbar_arrive(bar_a + 8 * c2);
          }
          if (tid == 0 && layer == 0 && c == 0) tc_stamp(dbg, h, 19);
        }
        if (tid == 0) tc_stamp(dbg, h, 5 + 2 * layer);
      }

      // output epilogue on this warpgroup's 16-column halves of D3
      mbar_wait(bar_d3, ph);
      tc_fence_after();
      if (tid == 0) tc_stamp(dbg, h, 8);
      float* sout = (states_out && valid) ? states_out + ((long long)h * R + row) * O : nullptr;
      uint8_t* xnext = xbuf + ((h + 1) & 1) * g.x_bytes;
      const int CC = max(g.Op >> 5, (SC + 3) >> 2);
#pragma unroll 1
      for (int cc = 0; cc < CC; ++cc) {
        const int col0 = 32 * cc + 16 * wg;
        uint32_t v[32];
        if (col0 < g.Op) {
          tmem_ld16(lane_base + col0, v);
          tmem_ld_wait();
          if (tid == 0 && cc == 0) tc_stamp(dbg, h, 10);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0u;
        }
        if (dbg && blockIdx.x == 0 && h == 0 && col0 < g.Op) {
#pragma unroll
          for (int i = 0; i < 16; ++i) dbg[(2 * kTcRows + trow) * kTcDbgCols + col0 + i] = __uint_as_float(v[i]);
        }
        // branch-free: padded table entries are zero, padded outputs are masked by select
        float y[16], term[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int o = col0 + i;        // o < Op always when col0 < Op; tables are Op long
          const int oc = o < g.Op ? o : 0;
          const float raw = __uint_as_float(v[i]);
          y[i] = o < O ? raw + t_b3[oc] : 0.f;              // normalised prediction == next input
          const float x = fmaf(raw, t_P[oc], t_Q[oc]);      // (s - goal) * w with s = y*sd + mu
          term[i] = o < O ? fast_sqrt(fmaf(x, x, m.alpha2)) - m.alpha : 0.f;
        }
        st_total += ((term[0] + term[1]) + (term[2] + term[3])) + ((term[4] + term[5]) + (term[6] + term[7])) +
                    (((term[8] + term[9]) + (term[10] + term[11])) + ((term[12] + term[13]) + (term[14] + term[15])));
        if (sout) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int o = col0 + i;
            if (o < O) sout[o] = fmaf(y[i], t_sd[o], t_mu[o]);  // unnormalize_state (data.py:255-257)
          }
        }
        if (h + 1 < H) {
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int j = 4 * cc + 2 * wg + jj;
            if (j < SC) {
              uint4 pk;
              pk.x = pack2<FP16>(y[8 * jj + 0], y[8 * jj + 1]); pk.y = pack2<FP16>(y[8 * jj + 2], y[8 * jj + 3]);
              pk.z = pack2<FP16>(y[8 * jj + 4], y[8 * jj + 5]); pk.w = pack2<FP16>(y[8 * jj + 6], y[8 * jj + 7]);
              *reinterpret_cast<uint4*>(xnext + (QA + j) * (kTcRows * 16) + trow * 16) = pk;
            }
          }
        }
      }
      if (tid == 0) tc_stamp(dbg, h, 11);
      if (h + 1 < H) {
        fence_proxy_async();   // generic-proxy writes of the input tile -> visible to the MMA
        tc_fence_before();     // our tcgen05.ld of D3 is ordered before the next layer-1 MMA
        if (tid == 0) tc_stamp(dbg, h, 14);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_x);
      }
      if (tid == 0) tc_stamp(dbg, h, 9);
    }
    costp[wg * kTcRows + trow] = st_total;
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid < kTcRows) {
    const long long row = (long long)blockIdx.x * kTcRows + tid;
    if (row < R) costs[row] = (costp[tid] + costp[kTcRows + tid]) + costp[2 * kTcRows + tid];
  }
  if (warp == kTcMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace mbrl
