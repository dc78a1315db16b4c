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

// out[j] = act(bias[j] + sum_k in[k] * Wt[k][j]); partial sums per K-slice in `part`.
template <bool RELU>
__device__ __forceinline__ void replay_layer(const float* __restrict__ Wt, const float* __restrict__ bias,
                                             const float* in, float* out, float* part, int K, int Nout) {
  const int Jp = (Nout + 31) & ~31;
  int nsl = kReplayThreads / Jp;
  if (nsl < 1) nsl = 1;
  if (nsl > 8) nsl = 8;
  const int ks = (K + nsl - 1) / nsl;
  for (int j0 = 0; j0 < Jp; j0 += kReplayThreads) {  // Nout > 1024 never happens (kMaxHidden)
    const int j = j0 + (threadIdx.x % Jp), sl = threadIdx.x / Jp;
    if (sl < nsl && j < Nout) {
      const int k0 = sl * ks, k1 = min(K, k0 + ks);
      float acc = 0.f;
#pragma unroll 8
      for (int k = k0; k < k1; ++k) acc = fmaf(in[k], __ldg(Wt + (long long)k * Nout + j), acc);
      part[sl * Jp + j] = acc;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < Nout; j += kReplayThreads) {
    float acc = __ldg(bias + j);
    for (int sl = 0; sl < nsl; ++sl) acc += part[sl * Jp + j];
    out[j] = RELU ? fmaxf(acc, 0.f) : acc;
  }
  __syncthreads();
}

// mu_hist/sd_hist: [I+1][E][H][A] (slot i = distribution sampled in iteration i).
// injected: [I][H*R][A] or null.  return_mean: emit mu_hist[iterations] instead.
__global__ void __launch_bounds__(kReplayThreads)
replay_kernel(ModelDev m, ActionSource src, Shape sh, const float* __restrict__ s0,
              const float* __restrict__ mu_hist, const float* __restrict__ sd_hist,
              const BestEver* __restrict__ best_ever, int iterations, int return_mean,
              float* __restrict__ out_states, float* __restrict__ out_actions,
              MbrlPlanInfo* __restrict__ info) {
  extern __shared__ __align__(16) float rs[];
  const int O = m.O, A = m.A, D = m.D, U = m.U, H = sh.H;
  const int Up = (U + 31) & ~31;
  float* acts = rs;              // [H][A]
  float* x = acts + H * A;       // [D]
  float* h1 = x + D;             // [U]
  float* h2 = h1 + U;            // [U]
  float* y = h2 + U;             // [O]
  float* part = y + O;           // [8][max(Up, Op)]
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
    replay_layer<true>(m.W1t, m.b1, x, h1, part, D, U);
    replay_layer<true>(m.W2t, m.b2, h1, h2, part, U, U);
    replay_layer<false>(m.W3t, m.b3, h2, x, part, U, O);  // x[0..O) <- normalised prediction
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
  (void)Up;
}

inline size_t replay_smem_bytes(int O, int A, int U, int H) {
  const int Up = (U + 31) & ~31, Op = (O + 31) & ~31;
  const int D = O + A;
  return sizeof(float) * (size_t)(H * A + D + 2 * U + O + 8 * (Up > Op ? Up : Op));
}

}  // namespace mbrl
