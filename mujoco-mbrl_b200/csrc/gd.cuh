// Batched GradientDescentPlanner (src/mbrl/planners.py:28-137).
//
// The reference optimises ONE action sequence [H, A] with Adam(lr 0.01): every iteration rolls the
// dynamics model forward H steps (planners.py:119-121), sums state_action_cost(s_{h+1}, a_h)
// (planners.py:123), back-propagates through the whole rollout (autograd) and takes an Adam step,
// stopping early when mean|a_old - a_new| < stop_condition (planners.py:128-135).  It returns the
// states of the LAST forward pass (computed with the actions before the final step) and the
// updated actions.  This kernel does exactly that for B independent restarts at once -- one CTA per
// restart, everything but the weights resident in shared memory, analytic back-propagation through
// time (no autograd graph):
//
//   forward   x = [norm(s_h), norm(a_h)];  h1 = relu(W1 x + b1);  h2 = relu(W2 h1 + b2);
//             s_{h+1} = unnorm(W3 h2 + b3)                       (models.py:13-29, 106-110; data.py:255-260)
//   backward  dL/ds_{h+1} = carried + SmoothAbs'(s_{h+1});  through W3^T, relu mask, W2^T, relu mask, W1^T;
//             dL/da_h = (W1^T g)[O:] / sd_a + (beta/A) sinh(a_h / beta);  carried = (W1^T g)[:O] / sd_s
//   Adam      torch.optim.Adam defaults (betas 0.9 / 0.999, eps 1e-8, bias correction), fp32
//
// fp32 CUDA cores throughout: batch-1 matrix-vector chains have nothing for tensor cores to do.  W2 (the
// bulk of the weights) is copied into shared memory when it fits next to the activations (hidden 200 at
// H = 30: 232,208 of 232,448 bytes), W1 / W3 (and W2 otherwise) are read from L2 (K-major fp32 copies of
// the handle).  A forward matvec splits K over the 16 warps (coalesced 128-byte rows, 16 partial sums
// reduced through shared memory), a transposed matvec gives every warp whole rows (coalesced, shuffle
// reduction).  Both issue their loads in batches of up to 8 independent rows / column blocks before the
// dependent FMA chain: the first version paid one L2 round trip per element (25 us per step) -- the order
// of the additions, and with it every result bit, is unchanged.
#pragma once
#include "common.cuh"

namespace mbrl {

constexpr int kGdThreads = 512;
constexpr int kGdWarps = kGdThreads / 32;

struct GdParams {
  int H, iterations;
  float lr, stop, beta1, beta2, eps;
};

struct GdLayout {
  int S, Aa, GA, M, V, X, H1, H2, part, gy, gh2, gh1, gx, gs, red, total;  // offsets in floats
  int Npad;
};

__host__ __device__ inline GdLayout gd_layout(int O, int A, int U, int H) {
  const int D = O + A;
  int widest = U > D ? U : D;
  widest = widest > O ? widest : O;
  GdLayout L;
  L.Npad = (widest + 31) & ~31;
  int off = 0;
  auto take = [&](int n) { const int o = off; off += (n + 3) & ~3; return o; };
  L.S = take((H + 1) * O); L.Aa = take(H * A); L.GA = take(H * A); L.M = take(H * A); L.V = take(H * A);
  L.X = take(H * D); L.H1 = take(H * U); L.H2 = take(H * U);
  L.part = take(kGdWarps * L.Npad);
  L.gy = take(O); L.gh2 = take(U); L.gh1 = take(U); L.gx = take(D); L.gs = take(O); L.red = take(64);
  L.total = off;
  return L;
}

template <bool GLOBAL>
__device__ __forceinline__ float gd_ld(const float* p) { return GLOBAL ? __ldg(p) : *p; }

// out[j] = act(bias[j] + sum_k Wt[k][j] x[k]),  Wt K-major [K][N].  GLOBAL: Wt in global memory (else shared).
template <bool GLOBAL>
__device__ __forceinline__ void gd_matvec_fwd(const float* __restrict__ Wt, const float* __restrict__ bias, const float* x,
                                              float* out, int K, int N, bool relu, float* part, int Npad) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int cs = 0; cs < N; cs += 256) {  // 8 column blocks of 32 at a time: 8 independent accumulators per lane
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
    for (int k = warp; k < K; k += kGdWarps) {
      const float xk = x[k];
      const float* row = Wt + (size_t)k * N + cs + lane;
      float w[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) w[c] = cs + 32 * c + lane < N ? gd_ld<GLOBAL>(row + 32 * c) : 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = fmaf(w[c], xk, acc[c]);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (cs + 32 * c < N) part[warp * Npad + cs + 32 * c + lane] = acc[c];  // Npad is a multiple of 32 >= N
  }
  __syncthreads();
  for (int j = threadIdx.x; j < N; j += kGdThreads) {
    float s = __ldg(bias + j);
#pragma unroll
    for (int w = 0; w < kGdWarps; ++w) s += part[w * Npad + j];
    out[j] = relu ? fmaxf(s, 0.f) : s;
  }
  __syncthreads();
}

// out[k] = (sum_j Wt[k][j] g[j]) * (mask ? mask[k] > 0 : 1)
template <bool GLOBAL>
__device__ __forceinline__ void gd_matvec_bwd(const float* __restrict__ Wt, const float* g, float* out, int K, int N,
                                              const float* mask) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 2
  for (int k = warp; k < K; k += kGdWarps) {
    float acc = 0.f;
    for (int js = 0; js < N; js += 256) {
      const float* row = Wt + (size_t)k * N + js + lane;
      float w[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) w[c] = js + 32 * c + lane < N ? gd_ld<GLOBAL>(row + 32 * c) : 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (js + 32 * c + lane < N) acc = fmaf(w[c], g[js + 32 * c + lane], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[k] = (mask == nullptr || mask[k] > 0.f) ? acc : 0.f;
  }
  __syncthreads();
}

// Register-resident variants for the two small matrices.  Both matvec directions touch element
// Wt[k][j] from the SAME thread -- warp = k mod 16, lane = j mod 32 -- so a thread loads its RI x RC
// elements (k = warp + 16 i, j = lane + 32 c; zero beyond K / N) once per plan and no step reads W1 or W3
// from memory again.  The additions run in the order of the memory-based functions above: same bits.
template <int RI, int RC>
__device__ __forceinline__ void gd_load_regs(const float* __restrict__ Wt, int K, int N, float (&w)[RI][RC]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < RI; ++i)
#pragma unroll
    for (int c = 0; c < RC; ++c) {
      const int k = warp + kGdWarps * i, j = lane + 32 * c;
      w[i][c] = (k < K && j < N) ? __ldg(Wt + (size_t)k * N + j) : 0.f;
    }
}
template <int RI, int RC>
__device__ __forceinline__ void gd_matvec_fwd_regs(const float (&w)[RI][RC], const float* __restrict__ bias, const float* x,
                                                   float* out, int K, int N, bool relu, float* part, int Npad) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[RC];
#pragma unroll
  for (int c = 0; c < RC; ++c) acc[c] = 0.f;
#pragma unroll
  for (int i = 0; i < RI; ++i) {
    const int k = warp + kGdWarps * i;
    if (k < K) {
      const float xk = x[k];
#pragma unroll
      for (int c = 0; c < RC; ++c) acc[c] = fmaf(w[i][c], xk, acc[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < RC; ++c)
    if (32 * c < N) part[warp * Npad + 32 * c + lane] = acc[c];
  __syncthreads();
  for (int j = threadIdx.x; j < N; j += kGdThreads) {
    float s = __ldg(bias + j);
#pragma unroll
    for (int ww = 0; ww < kGdWarps; ++ww) s += part[ww * Npad + j];
    out[j] = relu ? fmaxf(s, 0.f) : s;
  }
  __syncthreads();
}
template <int RI, int RC>
__device__ __forceinline__ void gd_matvec_bwd_regs(const float (&w)[RI][RC], const float* g, float* out, int K, int N,
                                                   const float* mask) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float gl[RC];
#pragma unroll
  for (int c = 0; c < RC; ++c) gl[c] = lane + 32 * c < N ? g[lane + 32 * c] : 0.f;
#pragma unroll
  for (int i = 0; i < RI; ++i) {
    const int k = warp + kGdWarps * i;
    if (k < K) {  // warp-uniform
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < RC; ++c)
        if (lane + 32 * c < N) acc = fmaf(w[i][c], gl[c], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) out[k] = (mask == nullptr || mask[k] > 0.f) ? acc : 0.f;
    }
  }
  __syncthreads();
}
// shapes whose W1 / W3 fit the register budget: D <= 32, hidden <= 208, O <= 32 (cartpole, cheetah, walker)
constexpr int kGdR1I = 2, kGdR1C = 7, kGdR3I = 13, kGdR3C = 1;
__host__ __device__ inline bool gd_small_in_regs(int O, int A, int U) {
  return O + A <= kGdWarps * kGdR1I && U <= 32 * kGdR1C && U <= kGdWarps * kGdR3I && O <= 32 * kGdR3C;
}

__device__ __forceinline__ float gd_block_sum(float v, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  if (warp == 0) {
    s = lane < kGdWarps ? red[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[32] = s;
  }
  __syncthreads();
  s = red[32];
  __syncthreads();
  return s;
}

// grid = restarts; s0 [O] is shared by all restarts; init_actions [B][H][A];
// out_states [B][H+1][O] (s_0 first, as the reference returns them), out_actions [B][H][A],
// out_cost [B] = the loss of the last forward pass, out_iters [B] = iterations run.
template <bool W2_SMEM, bool REGW>
__global__ void __launch_bounds__(kGdThreads, 1)
gd_plan_kernel(ModelDev m, GdParams gp, const float* __restrict__ s0, const float* __restrict__ init_actions,
               float* __restrict__ out_states, float* __restrict__ out_actions, float* __restrict__ out_cost,
               int* __restrict__ out_iters) {
  constexpr bool W2S = W2_SMEM;
  extern __shared__ __align__(16) float gd_smem[];
  const int O = m.O, A = m.A, D = m.D, U = m.U, H = gp.H;
  const GdLayout L = gd_layout(O, A, U, H);
  float *S = gd_smem + L.S, *Aa = gd_smem + L.Aa, *GA = gd_smem + L.GA, *M = gd_smem + L.M, *V = gd_smem + L.V;
  float *X = gd_smem + L.X, *H1 = gd_smem + L.H1, *H2 = gd_smem + L.H2, *part = gd_smem + L.part;
  float *gy = gd_smem + L.gy, *gh2 = gd_smem + L.gh2, *gh1 = gd_smem + L.gh1, *gx = gd_smem + L.gx, *gs = gd_smem + L.gs;
  float* red = gd_smem + L.red;
  const int tid = threadIdx.x, b = blockIdx.x;
  const int HA = H * A;

  const float* W2 = m.W2t;
  if (W2S) {  // 16-byte copies: U*U floats right after the activations (gd_layout keeps total a multiple of 4)
    float* w2s = gd_smem + L.total;
    const int n = U * U;
    if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(m.W2t) & 15) == 0) {
      for (int i = tid; i < n / 4; i += kGdThreads) reinterpret_cast<float4*>(w2s)[i] = __ldg(reinterpret_cast<const float4*>(m.W2t) + i);
    } else {
      for (int i = tid; i < n; i += kGdThreads) w2s[i] = __ldg(m.W2t + i);
    }
    W2 = w2s;
  }
  float w1r[REGW ? kGdR1I : 1][REGW ? kGdR1C : 1], w3r[REGW ? kGdR3I : 1][REGW ? kGdR3C : 1];
  if (REGW) { gd_load_regs(m.W1t, D, U, w1r); gd_load_regs(m.W3t, U, O, w3r); }
  for (int i = tid; i < O; i += kGdThreads) S[i] = s0[i];
  for (int i = tid; i < HA; i += kGdThreads) { Aa[i] = init_actions[(size_t)b * HA + i]; M[i] = 0.f; V[i] = 0.f; }
  __syncthreads();

  const float inv_beta = 1.0f / m.beta;
  float loss = 0.f;
  int ran = 0;
  float pow1 = 1.f, pow2 = 1.f;  // beta1^t, beta2^t
  for (int it = 0; it < gp.iterations; ++it) {
    // ---- forward: roll the model H steps, keep every layer's activations ----
    float cost_acc = 0.f;  // this thread's share of the loss
    for (int h = 0; h < H; ++h) {
      float* x = X + h * D;
      for (int d = tid; d < D; d += kGdThreads)
        x[d] = d < O ? __fdiv_rn(__fsub_rn(S[h * O + d], __ldg(m.mu_s + d)), __ldg(m.sd_s + d))
                     : __fdiv_rn(__fsub_rn(Aa[h * A + d - O], __ldg(m.mu_a + d - O)), __ldg(m.sd_a + d - O));
      __syncthreads();
      if (REGW) gd_matvec_fwd_regs(w1r, m.b1, x, H1 + h * U, D, U, true, part, L.Npad);
      else gd_matvec_fwd<true>(m.W1t, m.b1, x, H1 + h * U, D, U, true, part, L.Npad);
      gd_matvec_fwd<!W2S>(W2, m.b2, H1 + h * U, H2 + h * U, U, U, true, part, L.Npad);
      if (REGW) gd_matvec_fwd_regs(w3r, m.b3, H2 + h * U, gy, U, O, false, part, L.Npad);  // gy doubles as the y buffer
      else gd_matvec_fwd<true>(m.W3t, m.b3, H2 + h * U, gy, U, O, false, part, L.Npad);
      for (int o = tid; o < O; o += kGdThreads) {
        const float s = __fadd_rn(__fmul_rn(gy[o], __ldg(m.sd_s + o)), __ldg(m.mu_s + o));  // unnormalize_field
        S[(h + 1) * O + o] = s;
        cost_acc += smooth_abs_term(s, __ldg(m.goal + o), __ldg(m.cost_w + o), m.alpha, m.alpha2);
      }
      for (int a = tid; a < A; a += kGdThreads) cost_acc += m.beta2 * cosh_term(Aa[h * A + a], m.beta) / (float)A;
      __syncthreads();
    }
    loss = gd_block_sum(cost_acc, red);

    // ---- backward through time ----
    for (int o = tid; o < O; o += kGdThreads) gs[o] = 0.f;
    __syncthreads();
    for (int h = H - 1; h >= 0; --h) {
      for (int o = tid; o < O; o += kGdThreads) {
        // d/ds [sqrt(((s-g) w)^2 + alpha^2) - alpha] = (s-g) w^2 / sqrt(((s-g) w)^2 + alpha^2)   (models.py:255-259)
        const float w = __ldg(m.cost_w + o), xs = (S[(h + 1) * O + o] - __ldg(m.goal + o)) * w;
        const float dcost = xs * w / sqrtf(xs * xs + m.alpha2);
        gy[o] = (gs[o] + dcost) * __ldg(m.sd_s + o);
      }
      __syncthreads();
      if (REGW) gd_matvec_bwd_regs(w3r, gy, gh2, U, O, H2 + h * U);  // W3^T g, relu'(z2)
      else gd_matvec_bwd<true>(m.W3t, gy, gh2, U, O, H2 + h * U);
      gd_matvec_bwd<!W2S>(W2, gh2, gh1, U, U, H1 + h * U);           // W2^T g, relu'(z1)
      if (REGW) gd_matvec_bwd_regs(w1r, gh1, gx, D, U, nullptr);     // W1^T g
      else gd_matvec_bwd<true>(m.W1t, gh1, gx, D, U, nullptr);
      for (int d = tid; d < D; d += kGdThreads) {
        if (d < O) gs[d] = gx[d] / __ldg(m.sd_s + d);
        else {
          const int a = d - O;
          // d/da_j [beta^2 mean_a(cosh(a/beta) - 1)] = (beta / A) sinh(a_j / beta)   (models.py:271-272)
          GA[h * A + a] = gx[d] / __ldg(m.sd_a + a) + m.beta / (float)A * sinhf(Aa[h * A + a] * inv_beta);
        }
      }
      __syncthreads();
    }

    // ---- Adam step (torch.optim.Adam defaults) and the stop test (planners.py:128-135) ----
    pow1 *= gp.beta1; pow2 *= gp.beta2;
    const float step_size = gp.lr / (1.f - pow1), bc2_sqrt = sqrtf(1.f - pow2);
    float change = 0.f;
    for (int i = tid; i < HA; i += kGdThreads) {
      const float g = GA[i];
      const float mm = fmaf(g - M[i], 1.f - gp.beta1, M[i]);
      const float vv = fmaf(g * g, 1.f - gp.beta2, V[i] * gp.beta2);
      M[i] = mm; V[i] = vv;
      const float denom = sqrtf(vv) / bc2_sqrt + gp.eps;
      const float a_new = Aa[i] - step_size * (mm / denom);
      change += fabsf(Aa[i] - a_new);
      Aa[i] = a_new;
    }
    change = gd_block_sum(change, red);
    ++ran;
    if (change / (float)HA < gp.stop) break;
  }

  for (int i = tid; i < (H + 1) * O; i += kGdThreads) out_states[(size_t)b * (H + 1) * O + i] = S[i];
  for (int i = tid; i < HA; i += kGdThreads) out_actions[(size_t)b * HA + i] = Aa[i];
  if (tid == 0) { out_cost[b] = loss; out_iters[b] = ran; }
}

}  // namespace mbrl
