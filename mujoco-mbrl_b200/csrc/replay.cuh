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

constexpr int kReplayThreads = 1024;
constexpr int kReplayMaxSlices = 16;

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
    float acc = __ldg(bias + j);
    for (int sl = 0; sl < sp.slices; ++sl) acc += part[sl * sp.ld + j];
    out[j] = RELU ? fmaxf(acc, 0.f) : acc;
  }
  __syncthreads();
}

struct ReplayLayout {  // shared-memory carve-up, in floats
  int acts, x, h1, h2, y, part, w1, w2, w3, total;
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
  L.w1 = L.part + kReplayMaxSlices * (ldu > ldo ? ldu : ldo);
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
              const BestEver* __restrict__ best_ever, int iterations, int return_mean,
              float* __restrict__ out_states, float* __restrict__ out_actions,
              MbrlPlanInfo* __restrict__ info) {
  extern __shared__ __align__(16) float rs[];
  const int O = m.O, A = m.A, D = m.D, U = m.U, H = sh.H;
  const ReplayLayout L = replay_layout(O, A, U, H, SMEM_W);
  float *acts = rs + L.acts, *x = rs + L.x, *h1 = rs + L.h1, *h2 = rs + L.h2, *y = rs + L.y, *part = rs + L.part;
  const ReplaySplit sp1 = replay_split(D, U), sp2 = replay_split(U, U), sp3 = replay_split(U, O);
  const float *W1 = m.W1t, *W2 = m.W2t, *W3 = m.W3t;
  if (SMEM_W) {
    float *w1 = rs + L.w1, *w2 = rs + L.w2, *w3 = rs + L.w3;
    for (int i = threadIdx.x; i < D * sp1.ld; i += kReplayThreads) {
      const int k = i / sp1.ld, j = i - k * sp1.ld;
      w1[i] = j < U ? __ldg(m.W1t + (long long)k * U + j) : 0.f;
    }
    for (int i = threadIdx.x; i < U * sp2.ld; i += kReplayThreads) {
      const int k = i / sp2.ld, j = i - k * sp2.ld;
      w2[i] = j < U ? __ldg(m.W2t + (long long)k * U + j) : 0.f;
    }
    for (int i = threadIdx.x; i < U * sp3.ld; i += kReplayThreads) {
      const int k = i / sp3.ld, j = i - k * sp3.ld;
      w3[i] = j < O ? __ldg(m.W3t + (long long)k * O + j) : 0.f;
    }
    W1 = w1; W2 = w2; W3 = w3;
  }
  const int env_l = blockIdx.x;
  const long long R = sh.rows();
  const long long EHA = (long long)sh.E * H * A;

  const BestEver b = best_ever[env_l];
  if (return_mean) {
    const float* mu = mu_hist + (long long)iterations * EHA + (long long)env_l * H * A;
    for (int i = threadIdx.x; i < H * A; i += kReplayThreads) acts[i] = clipf(mu[i], src.lo, src.hi);
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
  for (int o = threadIdx.x; o < O; o += kReplayThreads) y[o] = __ldg(s0 + (long long)env_l * O + o);
  __syncthreads();

  for (int h = 0; h < H; ++h) {
    for (int i = threadIdx.x; i < D; i += kReplayThreads) {
      x[i] = i < O ? __fdiv_rn(__fsub_rn(y[i], __ldg(m.mu_s + i)), __ldg(m.sd_s + i))
                   : __fdiv_rn(__fsub_rn(acts[h * A + i - O], __ldg(m.mu_a + i - O)), __ldg(m.sd_a + i - O));
    }
    __syncthreads();
    replay_layer<true, SMEM_W>(W1, m.b1, x, h1, part, D, U, sp1);
    replay_layer<true, SMEM_W>(W2, m.b2, h1, h2, part, U, U, sp2);
    replay_layer<false, SMEM_W>(W3, m.b3, h2, x, part, U, O, sp3);  // x[0..O) <- normalised prediction
    for (int o = threadIdx.x; o < O; o += kReplayThreads) {
      const float s = __fadd_rn(__fmul_rn(x[o], __ldg(m.sd_s + o)), __ldg(m.mu_s + o));
      y[o] = s;
      out_states[((long long)env_l * H + h) * O + o] = s;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < H * A; i += kReplayThreads) out_actions[(long long)env_l * H * A + i] = acts[i];
  if (info && threadIdx.x == 0) {
    info[env_l].best_cost = b.cost;
    info[env_l].best_iteration = b.iteration;
    info[env_l].best_index = b.index;
    info[env_l].reserved = 0;
  }
}

}  // namespace mbrl
