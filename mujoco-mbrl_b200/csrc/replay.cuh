// Emits the chosen plan: regenerates the winning candidate's action sequence from the
// sampler (or the final mean sequence) and replays it through the fp32 dynamics model to
// produce the predicted states s_1..s_H that RandomShootingPlanner.plan returns
// (src/mbrl/planners.py:184-187, 212-215).  One CTA per environment; a batch-1, H-step
// recurrence is latency-bound, so the layer is K-sliced over the whole CTA.
#pragma once
#include "common.cuh"
#include "philox.cuh"
#include "select.cuh"

namespace mbrl {

constexpr int kReplayThreads = 256;  // measured: 256 beats 512 and 1024 (block barriers dominate; the layer is smem-bandwidth bound)
constexpr int kReplayMaxSlices = 16;
#ifdef MBRL_REPLAY_PROFILE
__device__ long long g_replay_stamps[16];
#define REPLAY_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_replay_stamps[i] = clock64(); } while (0)
#else
#define REPLAY_STAMP(i) do { } while (0)
#endif

// Work split of one dense layer over the CTA: thread -> (output quad q, K-slice s).  Every
// thread accumulates 4 adjacent outputs over its K-slice (one 16-byte weight load feeds 4 FMAs),
// partial sums meet in shared memory.
struct ReplaySplit {
  int quads, slices, ks, ld;  // ld = padded row stride of the weight / partial arrays (multiple of 4)
  int q, s;                   // this thread's quad / slice (s >= slices: idle)
};
__device__ __forceinline__ ReplaySplit replay_split(int K, int Nout) {
  ReplaySplit r;
  r.quads = (Nout + 3) >> 2;
  r.ld = r.quads * 4;
  r.slices = min(min(kReplayThreads / r.quads, K), kReplayMaxSlices);
  r.ks = (K + r.slices - 1) / r.slices;
  r.q = threadIdx.x % r.quads;
  r.s = threadIdx.x / r.quads;
  return r;
}

// out[j] = act(bias[j] + sum_k in[k] * W[k][j]).  SMEM_W: W is the padded shared-memory copy
// (row stride sp.ld, 16-byte aligned); otherwise the K-major global array with row stride Nout.
template <bool RELU, bool SMEM_W>
__device__ __forceinline__ void replay_layer(const float* __restrict__ W, const float* __restrict__ bias,
                                             const float* in, float* out, float* part, int K, int Nout,
                                             const ReplaySplit& sp) {
  if (sp.s < sp.slices) {
    const int k0 = sp.s * sp.ks, k1 = min(K, k0 + sp.ks);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (SMEM_W) {
      const float* w = W + 4 * sp.q;
#pragma unroll 4
      for (int k = k0; k < k1; ++k) {
        const float4 w4 = *reinterpret_cast<const float4*>(w + (long long)k * sp.ld);
        const float a = in[k];
        acc.x = fmaf(a, w4.x, acc.x); acc.y = fmaf(a, w4.y, acc.y);
        acc.z = fmaf(a, w4.z, acc.z); acc.w = fmaf(a, w4.w, acc.w);
      }
    } else {
      const int j = 4 * sp.q;
#pragma unroll 4
      for (int k = k0; k < k1; ++k) {
        const float* w = W + (long long)k * Nout + j;
        const float a = in[k];
        acc.x = fmaf(a, __ldg(w), acc.x);
        if (j + 1 < Nout) acc.y = fmaf(a, __ldg(w + 1), acc.y);
        if (j + 2 < Nout) acc.z = fmaf(a, __ldg(w + 2), acc.z);
        if (j + 3 < Nout) acc.w = fmaf(a, __ldg(w + 3), acc.w);
      }
    }
    *reinterpret_cast<float4*>(part + sp.s * sp.ld + 4 * sp.q) = acc;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < Nout; j += kReplayThreads) {
    float acc = bias[j];
    for (int sl = 0; sl < sp.slices; ++sl) acc += part[sl * sp.ld + j];
    out[j] = RELU ? fmaxf(acc, 0.f) : acc;
  }
  __syncthreads();
}

struct ReplayLayout {  // shared-memory carve-up, in floats
  int acts, x, h1, h2, y, part, vec, w1, w2, w3, total;  // vec: b1,b2 [ldu each], b3,mu_s,sd_s [ldo each], mu_a,sd_a [32 each]
};
__host__ __device__ inline ReplayLayout replay_layout(int O, int A, int U, int H, bool smem_w) {
  auto r4 = [](int v) { return (v + 3) & ~3; };
  ReplayLayout L;
  const int D = O + A, ldu = r4(U), ldo = r4(O);
  L.acts = 0;
  L.x = r4(H * A);
  L.h1 = L.x + r4(D);
  L.h2 = L.h1 + ldu;
  L.y = L.h2 + ldu;
  L.part = L.y + ldo;
  L.vec = L.part + kReplayMaxSlices * (ldu > ldo ? ldu : ldo);
  L.w1 = L.vec + 2 * ldu + 3 * ldo + 2 * kMaxAct;
  L.w2 = L.w1 + (smem_w ? D * ldu : 0);
  L.w3 = L.w2 + (smem_w ? U * ldu : 0);
  L.total = L.w3 + (smem_w ? U * ldo : 0);
  return L;
}
inline size_t replay_smem_bytes(int O, int A, int U, int H, bool smem_weights) {
  return sizeof(float) * (size_t)replay_layout(O, A, U, H, smem_weights).total;
}

// mu_hist/sd_hist: [I+1][E][H][A] (slot i = distribution sampled in iteration i).
// injected: [I][H*R][A] or null.  return_mean: emit mu_hist[iterations] instead.
// SMEM_W: the three fp32 weight matrices are copied into shared memory once (they are re-read
// every step of the recurrence; from L2 that costs ~3 dependent global-load latencies a step).
template <bool SMEM_W>
__global__ void __launch_bounds__(kReplayThreads)
replay_kernel(ModelDev m, ActionSource src, Shape sh, const float* __restrict__ s0,
              const float* __restrict__ mu_hist, const float* __restrict__ sd_hist,
              const BestEver* __restrict__ best_ever, int iterations, int return_mean, int actions_only,
              float* __restrict__ out_states, float* __restrict__ out_actions,
              MbrlPlanInfo* __restrict__ info) {
  extern __shared__ __align__(16) float rs[];
  const int O = m.O, A = m.A, D = m.D, U = m.U, H = sh.H;
  REPLAY_STAMP(0);
  const ReplayLayout L = replay_layout(O, A, U, H, SMEM_W);
  float *acts = rs + L.acts, *x = rs + L.x, *h1 = rs + L.h1, *h2 = rs + L.h2, *y = rs + L.y, *part = rs + L.part;
  const ReplaySplit sp1 = replay_split(D, U), sp2 = replay_split(U, U), sp3 = replay_split(U, O);
  // small per-model vectors live in shared memory: every step of the recurrence re-reads them
  float *vb1 = rs + L.vec, *vb2 = vb1 + sp1.ld, *vb3 = vb2 + sp2.ld, *vmu = vb3 + sp3.ld, *vsd = vmu + sp3.ld;
  float *vmua = vsd + sp3.ld, *vsda = vmua + kMaxAct;
  for (int i = threadIdx.x; i < U; i += kReplayThreads) { vb1[i] = __ldg(m.b1 + i); vb2[i] = __ldg(m.b2 + i); }
  for (int i = threadIdx.x; i < O; i += kReplayThreads) { vb3[i] = __ldg(m.b3 + i); vmu[i] = __ldg(m.mu_s + i); vsd[i] = __ldg(m.sd_s + i); }
  for (int i = threadIdx.x; i < A; i += kReplayThreads) { vmua[i] = __ldg(m.mu_a + i); vsda[i] = __ldg(m.sd_a + i); }
  const float *W1 = m.W1t, *W2 = m.W2t, *W3 = m.W3t;
  if (SMEM_W && !actions_only) {
    float *w1 = rs + L.w1, *w2 = rs + L.w2, *w3 = rs + L.w3;
    auto copy_padded = [&](float* dst, const float* src, int rows, int cols, int ld) {
      if (ld == cols && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {  // contiguous: 16-byte copies
        const int n4 = rows * cols / 4;
#pragma unroll 8
        for (int i = threadIdx.x; i < n4; i += kReplayThreads)
          reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
      } else {
        for (int i = threadIdx.x; i < rows * ld; i += kReplayThreads) {
          const int k = i / ld, j = i - k * ld;
          dst[i] = j < cols ? __ldg(src + (long long)k * cols + j) : 0.f;
        }
      }
    };
    copy_padded(w1, m.W1t, D, U, sp1.ld);
    copy_padded(w2, m.W2t, U, U, sp2.ld);
    copy_padded(w3, m.W3t, U, O, sp3.ld);
    W1 = w1; W2 = w2; W3 = w3;
  }
  const int env_l = blockIdx.x;
  const long long R = sh.rows();
  const long long EHA = (long long)sh.E * H * A;

  // everything above only touched the (static) model: under PDL it overlaps the last top-k
  pdl_trigger();
  pdl_wait();
  BestEver b;
  {
    const int4 raw = dep_load(reinterpret_cast<const int4*>(best_ever + env_l));  // written by the last top-k
    b.cost = __int_as_float(raw.x); b.iteration = raw.y; b.index = raw.z; b.pad = raw.w;
  }
  if (return_mean) {
    const float* mu = mu_hist + (long long)iterations * EHA + (long long)env_l * H * A;
    for (int i = threadIdx.x; i < H * A; i += kReplayThreads) acts[i] = clipf(dep_load(mu + i), src.lo, src.hi);
  } else {
    ActionSource s = src;
    s.iteration = (uint32_t)b.iteration;
    s.mu = mu_hist + (long long)b.iteration * EHA;
    s.sd = sd_hist + (long long)b.iteration * EHA;
    if (s.buf) s.buf += (long long)b.iteration * H * R * A;
    const long long row = (long long)env_l * sh.N + b.index;
    for (int h = threadIdx.x; h < H; h += kReplayThreads)
      for_each_action(s, A, H, h, env_l, b.index, row, R, [&](int a, float v) { acts[h * A + a] = v; });
  }
  for (int o = threadIdx.x; o < O; o += kReplayThreads) y[o] = dep_load(s0 + (long long)env_l * O + o);
  __syncthreads();
  REPLAY_STAMP(1);

  if (actions_only) {
    for (int i = threadIdx.x; i < H * O; i += kReplayThreads) out_states[(long long)env_l * H * O + i] = 0.f;
  }
  for (int h = 0; h < (actions_only ? 0 : H); ++h) {
    for (int i = threadIdx.x; i < D; i += kReplayThreads) {
      x[i] = i < O ? __fdiv_rn(__fsub_rn(y[i], vmu[i]), vsd[i])
                   : __fdiv_rn(__fsub_rn(acts[h * A + i - O], vmua[i - O]), vsda[i - O]);
    }
    __syncthreads();
    if (h == 5) REPLAY_STAMP(2);
    replay_layer<true, SMEM_W>(W1, vb1, x, h1, part, D, U, sp1);
    if (h == 5) REPLAY_STAMP(3);
    replay_layer<true, SMEM_W>(W2, vb2, h1, h2, part, U, U, sp2);
    if (h == 5) REPLAY_STAMP(4);
    replay_layer<false, SMEM_W>(W3, vb3, h2, x, part, U, O, sp3);  // x[0..O) <- normalised prediction
    if (h == 5) REPLAY_STAMP(5);
    for (int o = threadIdx.x; o < O; o += kReplayThreads) {
      const float s = __fadd_rn(__fmul_rn(x[o], vsd[o]), vmu[o]);
      y[o] = s;
      out_states[((long long)env_l * H + h) * O + o] = s;
    }
    __syncthreads();
    if (h == 5) REPLAY_STAMP(6);
  }
  REPLAY_STAMP(7);
  for (int i = threadIdx.x; i < H * A; i += kReplayThreads) out_actions[(long long)env_l * H * A + i] = acts[i];
  if (info && threadIdx.x == 0) {
    info[env_l].best_cost = b.cost;
    info[env_l].best_iteration = b.iteration;
    info[env_l].best_index = b.index;
    info[env_l].reserved = 0;
  }
}

// --------------------------------------------------------------------------------------
// Register-resident variant.  The recurrence is one environment, batch 1: per step the only
// real work is streaming W2 (U*U floats) past the hidden vector; everything else is a chain of
// dependent shared-memory latencies and block barriers, and barriers get dearer with the thread
// count.  So: 256 threads, and W2 lives in their REGISTERS -- thread (s, q) holds the 4 output
// columns 4q..4q+3 of the K slice [s*KPT2, s*KPT2+KPT2) (160 floats for U = 200).  W1 and W3
// (small) sit in shared memory, W3 transposed and row-skewed so that 8 lanes reduce one output
// with conflict-free loads + 3 shuffles.  Trip counts are template constants and every buffer is
// zero-padded to them: straight-line code, independent loads, no per-element predicates.
// Layer 1 (K = D, short) is one thread per output.  4 block barriers per step.  Shapes that fit
// no instantiation use replay_kernel.
// --------------------------------------------------------------------------------------
constexpr int kRegThreads = 256;

struct RegGeom {
  int S, kpt2, Q, ldu, ldo, kp1, kp2, k3, ld3, o3;
  int acts, x, y, h1, h2, part, vec, w1, w3, total;  // shared-memory offsets, in floats
  bool ok;
};
__host__ __device__ inline RegGeom replay_reg_geometry(int O, int A, int U, int H) {
  auto r4 = [](int v) { return (v + 3) & ~3; };
  RegGeom g;
  const int D = O + A;
  if (U <= 128) { g.S = 8; g.kpt2 = 16; } else { g.S = 5; g.kpt2 = 40; }
  g.Q = (U + 3) / 4;
  g.ok = U >= 1 && U <= g.S * g.kpt2 && g.Q * g.S <= kRegThreads && O <= 128 && A <= kMaxAct;
  g.ldu = 4 * g.Q;
  g.ldo = r4(O);
  g.kp1 = (D + 7) & ~7;
  g.kp2 = g.S * g.kpt2;
  g.k3 = (g.kp2 + 7) / 8;  // layer-3 K elements per lane (8 lanes per output)
  g.ld3 = 8 * g.k3 + 8;    // +8 words: the 4 outputs of a warp land in 4 distinct bank octets
  g.o3 = r4(O);            // outputs padded to whole warps of 4
  g.acts = 0;
  g.x = r4(H * A);
  g.y = g.x + g.kp1;
  g.h1 = g.y + g.ldo;
  g.h2 = g.h1 + g.kp2;
  g.part = g.h2 + 8 * g.k3;
  g.vec = g.part + g.S * g.ldu;
  g.w1 = g.vec + 2 * g.ldu + 3 * g.ldo + 2 * kMaxAct;
  g.w3 = g.w1 + g.kp1 * g.ldu;
  g.total = g.w3 + g.o3 * g.ld3;
  return g;
}

template <int S, int KPT2>
__global__ void __launch_bounds__(kRegThreads, 1)
replay_reg_kernel(ModelDev m, ActionSource src, Shape sh, RegGeom g, const float* __restrict__ s0,
                  const float* __restrict__ mu_hist, const float* __restrict__ sd_hist,
                  const BestEver* __restrict__ best_ever, int iterations, int return_mean,
                  float* __restrict__ out_states, float* __restrict__ out_actions,
                  MbrlPlanInfo* __restrict__ info) {
  extern __shared__ __align__(16) float rr[];
  constexpr int K3 = (S * KPT2 + 7) / 8;
  const int O = m.O, A = m.A, D = m.D, U = m.U, H = sh.H;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  REPLAY_STAMP(0);
  float *acts = rr + g.acts, *x = rr + g.x, *y = rr + g.y, *h1 = rr + g.h1, *h2 = rr + g.h2, *part = rr + g.part;
  float *vb1 = rr + g.vec, *vb2 = vb1 + g.ldu, *vb3 = vb2 + g.ldu, *vmu = vb3 + g.ldo, *vsd = vmu + g.ldo;
  float *vmua = vsd + g.ldo, *vsda = vmua + kMaxAct;
  float *w1 = rr + g.w1, *w3 = rr + g.w3;
  const bool active = t < S * g.Q;
  const int q = active ? t % g.Q : 0, s = active ? t / g.Q : 0;

  // ---- static model: overlaps the previous kernel under PDL ----
  float w2[KPT2][4];
#pragma unroll
  for (int i = 0; i < KPT2; ++i) {
    const int k = s * KPT2 + i;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      w2[i][c] = (active && k < U && 4 * q + c < U) ? __ldg(m.W2t + (long long)k * U + 4 * q + c) : 0.f;
  }
  for (int i = t; i < U; i += kRegThreads) { vb1[i] = __ldg(m.b1 + i); vb2[i] = __ldg(m.b2 + i); }
  for (int i = t; i < O; i += kRegThreads) { vb3[i] = __ldg(m.b3 + i); vmu[i] = __ldg(m.mu_s + i); vsd[i] = __ldg(m.sd_s + i); }
  for (int i = t; i < A; i += kRegThreads) { vmua[i] = __ldg(m.mu_a + i); vsda[i] = __ldg(m.sd_a + i); }
  for (int i = t; i < g.kp1 * g.ldu; i += kRegThreads) {
    const int k = i / g.ldu, c = i - k * g.ldu;
    w1[i] = (k < D && c < U) ? __ldg(m.W1t + (long long)k * U + c) : 0.f;
  }
  for (int i = t; i < g.o3 * g.ld3; i += kRegThreads) {  // transposed: w3[o][k]
    const int o = i / g.ld3, k = i - o * g.ld3;
    w3[i] = (o < O && k < U) ? __ldg(m.W3t + (long long)k * O + o) : 0.f;
  }
  for (int i = t; i < g.kp1; i += kRegThreads) x[i] = 0.f;
  for (int i = t; i < g.kp2; i += kRegThreads) h1[i] = 0.f;
  for (int i = t; i < 8 * g.k3; i += kRegThreads) h2[i] = 0.f;

  const int env_l = blockIdx.x;
  pdl_trigger();
  pdl_wait();
  BestEver b;
  {
    const int4 raw = dep_load(reinterpret_cast<const int4*>(best_ever + env_l));  // written by the last top-k
    b.cost = __int_as_float(raw.x); b.iteration = raw.y; b.index = raw.z; b.pad = raw.w;
  }
  {
    const long long R = sh.rows();
    const long long EHA = (long long)sh.E * H * A;
    if (return_mean) {
      const float* mu = mu_hist + (long long)iterations * EHA + (long long)env_l * H * A;
      for (int i = t; i < H * A; i += kRegThreads) acts[i] = clipf(dep_load(mu + i), src.lo, src.hi);
    } else {
      ActionSource as = src;
      as.iteration = (uint32_t)b.iteration;
      as.mu = mu_hist + (long long)b.iteration * EHA;
      as.sd = sd_hist + (long long)b.iteration * EHA;
      if (as.buf) as.buf += (long long)b.iteration * H * R * A;
      const long long row = (long long)env_l * sh.N + b.index;
      for (int h = t; h < H; h += kRegThreads)
        for_each_action(as, A, H, h, env_l, b.index, row, R, [&](int a, float v) { acts[h * A + a] = v; });
    }
  }
  __syncthreads();
  for (int i = t; i < D; i += kRegThreads) {
    x[i] = i < O ? __fdiv_rn(__fsub_rn(dep_load(s0 + (long long)env_l * O + i), vmu[i]), vsd[i])
                 : __fdiv_rn(__fsub_rn(acts[i - O], vmua[i - O]), vsda[i - O]);
  }
  __syncthreads();
  REPLAY_STAMP(1);

  const float* w1_t = w1 + t;
  const float4* h1_sl = reinterpret_cast<const float4*>(h1 + s * KPT2);
  float4* part_w = reinterpret_cast<float4*>(part + s * g.ldu + 4 * q);
  const float* part_r = part + t;
  const int o3 = warp * 4 + (lane >> 3), kl = lane & 7;  // layer 3: 4 outputs per warp, 8 lanes each
  const float* w3_l = w3 + o3 * g.ld3 + kl;
  const float* h2_l = h2 + kl;
  const bool l3 = o3 < g.o3;
  float* out_env = out_states + (long long)env_l * H * O;

  for (int h = 0; h < H; ++h) {
    if (h == 5) REPLAY_STAMP(2);
    if (t < U) {
      float a0 = vb1[t], a1 = 0.f;
      for (int k = 0; k < g.kp1; k += 8) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          a0 = fmaf(x[k + i], w1_t[(k + i) * g.ldu], a0);
          a1 = fmaf(x[k + i + 1], w1_t[(k + i + 1) * g.ldu], a1);
        }
      }
      h1[t] = fmaxf(a0 + a1, 0.f);
    }
    __syncthreads();
    if (h == 5) REPLAY_STAMP(3);
    // layer 2: W2 from registers
    if (active) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i4 = 0; i4 < KPT2 / 4; ++i4) {
        const float4 a = h1_sl[i4];
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc.x = fmaf(av[e], w2[i4 * 4 + e][0], acc.x);
          acc.y = fmaf(av[e], w2[i4 * 4 + e][1], acc.y);
          acc.z = fmaf(av[e], w2[i4 * 4 + e][2], acc.z);
          acc.w = fmaf(av[e], w2[i4 * 4 + e][3], acc.w);
        }
      }
      *part_w = acc;
    }
    if (h == 5) REPLAY_STAMP(8);
    __syncthreads();
    if (h == 5) REPLAY_STAMP(9);
    if (t < U) {
      float p[S];
#pragma unroll
      for (int sl = 0; sl < S; ++sl) p[sl] = part_r[sl * g.ldu];
      float acc = vb2[t];
#pragma unroll
      for (int sl = 0; sl < S; ++sl) acc += p[sl];
      h2[t] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    if (h == 5) REPLAY_STAMP(4);
    // layer 3: 8 lanes per output
    if (l3) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int i = 0; i < K3; i += 2) {
        a0 = fmaf(h2_l[8 * i], w3_l[8 * i], a0);
        if (i + 1 < K3) a1 = fmaf(h2_l[8 * (i + 1)], w3_l[8 * (i + 1)], a1);
      }
      float acc = a0 + a1;
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (kl == 0 && o3 < O) {
        const float sv = __fadd_rn(__fmul_rn(acc + vb3[o3], vsd[o3]), vmu[o3]);  // unnormalize_state (data.py:255-257)
        y[o3] = sv;
        out_env[h * O + o3] = sv;
        x[o3] = __fdiv_rn(__fsub_rn(sv, vmu[o3]), vsd[o3]);                      // next step's normalize_state
      }
    }
    if (t >= kRegThreads - A && h + 1 < H) {
      const int a = t - (kRegThreads - A);
      x[O + a] = __fdiv_rn(__fsub_rn(acts[(h + 1) * A + a], vmua[a]), vsda[a]);
    }
    __syncthreads();
    if (h == 5) REPLAY_STAMP(5);
  }
  REPLAY_STAMP(7);
  for (int i = t; i < H * A; i += kRegThreads) out_actions[(long long)env_l * H * A + i] = acts[i];
  if (info && t == 0) {
    info[env_l].best_cost = b.cost;
    info[env_l].best_iteration = b.iteration;
    info[env_l].best_index = b.index;
    info[env_l].reserved = 0;
  }
}

}  // namespace mbrl
