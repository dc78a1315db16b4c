// fp32 CUDA-core rollout + cost kernel (MBRL_ENGINE_SIMT_FP32).
//
// Fuses the hot loop of RandomShootingPlanner._generate_trajectories
// (src/mbrl/planners.py:199-210) with DynamicsModel.forward (src/mbrl/models.py:13-29),
// Model._forward (models.py:106-110), the normalisers (src/mbrl/data.py:255-260) and
// state_action_cost (src/mbrl/agents.py:182-183, models.py:244-272).  All arithmetic is
// fp32 with the reference's operation order inside a row (separate mul/add where torch
// rounds twice, true division by std), so per-step states agree with the CPU planner to
// ~1e-6 relative; only the GEMM summation order differs from MKL.
//
// This is the parity engine and the cross-check for the tcgen05 engine; it is not the
// throughput engine.  A CTA owns TM candidate rows for all H steps; activations stay in
// shared memory ([k][TM], rows contiguous so a warp's 8-row strip is two LDS.128 broadcast
// loads); weights are read through L1/L2 from a K-major copy (coalesced over outputs).
// Only costs[R] leave the SM (plus optional debug trajectories).
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace mbrl {

// out[j][r] = act(bias[j] + sum_k in[k][r] * Wt[k][j]) for r in this warp's 8-row strip.
// Warp (rg, cg): rows rg*8..rg*8+7 (rg < TM/8), column group cg of CG -- the CG warps of a strip
// split the output columns (passes of 32*CPT columns, interleaved over cg), so a wide layer keeps
// CG times more warps in flight on the same shared-memory tile; lane: columns c0 + c*32 + lane, c < CPT.
template <int TM, int CPT, int CG, bool RELU>
__device__ __forceinline__ void dense_layer(const float* __restrict__ Wt, const float* __restrict__ bias,
                                            const float* __restrict__ in, float* __restrict__ out,
                                            int K, int Nout) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rg = warp % (TM / 8), cg = warp / (TM / 8);
  const float* in_strip = in + rg * 8;
  for (int c0 = cg * 32 * CPT; c0 < Nout; c0 += CG * 32 * CPT) {
    float acc[CPT][8];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int j = c0 + c * 32 + lane;
      const float b = (j < Nout) ? __ldg(bias + j) : 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[c][i] = b;
    }
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(in_strip + k * TM);
      const float4 a1 = *reinterpret_cast<const float4*>(in_strip + k * TM + 4);
      const float* wrow = Wt + (long long)k * Nout + c0 + lane;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const float w = (c0 + c * 32 + lane < Nout) ? __ldg(wrow + c * 32) : 0.0f;
        acc[c][0] = fmaf(a0.x, w, acc[c][0]);
        acc[c][1] = fmaf(a0.y, w, acc[c][1]);
        acc[c][2] = fmaf(a0.z, w, acc[c][2]);
        acc[c][3] = fmaf(a0.w, w, acc[c][3]);
        acc[c][4] = fmaf(a1.x, w, acc[c][4]);
        acc[c][5] = fmaf(a1.y, w, acc[c][5]);
        acc[c][6] = fmaf(a1.z, w, acc[c][6]);
        acc[c][7] = fmaf(a1.w, w, acc[c][7]);
      }
    }
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int j = c0 + c * 32 + lane;
      if (j < Nout) {
        float4 o0, o1;
        o0.x = RELU ? fmaxf(acc[c][0], 0.f) : acc[c][0];
        o0.y = RELU ? fmaxf(acc[c][1], 0.f) : acc[c][1];
        o0.z = RELU ? fmaxf(acc[c][2], 0.f) : acc[c][2];
        o0.w = RELU ? fmaxf(acc[c][3], 0.f) : acc[c][3];
        o1.x = RELU ? fmaxf(acc[c][4], 0.f) : acc[c][4];
        o1.y = RELU ? fmaxf(acc[c][5], 0.f) : acc[c][5];
        o1.z = RELU ? fmaxf(acc[c][6], 0.f) : acc[c][6];
        o1.w = RELU ? fmaxf(acc[c][7], 0.f) : acc[c][7];
        *reinterpret_cast<float4*>(out + j * TM + rg * 8) = o0;
        *reinterpret_cast<float4*>(out + j * TM + rg * 8 + 4) = o1;
      }
    }
  }
}

// Shared memory: bufA [KA][TM] (layer-1 input, then layer-2 output), bufB [U][TM]
// (layer-1 output, then layer-3 output), bufS [O][TM] current un-normalised state, bufN [A][TM]
// the step's normalised actions (kept for the reward-head cost's second trunk evaluation).
template <int TM>
__host__ __device__ inline size_t simt_smem_bytes(int O, int A, int U) {
  const int KA = (O + A) > U ? (O + A) : U;
  const int KB = U > O ? U : O;
  return sizeof(float) * (size_t)TM * (KA + KB + O + A);
}

template <int TM, int CPT, int CG>
__global__ void __launch_bounds__(TM * 4 * CG)
rollout_simt_kernel(ModelDev m, ActionSource src, Shape sh, const float* __restrict__ s0,
                    float* __restrict__ costs, float* __restrict__ states_out,
                    float* __restrict__ actions_out) {
  extern __shared__ __align__(16) float simt_smem[];
  float* const smem = simt_smem;
  pdl_trigger();
  pdl_wait();
  const int O = m.O, A = m.A, D = m.D, U = m.U;
  const int KA = D > U ? D : U;
  const int KB = U > O ? U : O;
  float* bufA = smem;
  float* bufB = bufA + KA * TM;
  float* bufS = bufB + KB * TM;
  float* bufN = bufS + O * TM;
  const bool reward_head = m.cost_kind == MBRL_COST_REWARD_HEAD;

  const long long R = sh.rows();
  const int t = threadIdx.x;
  const long long row = (long long)blockIdx.x * TM + t;  // meaningful for t < TM
  const bool row_thread = t < TM;
  const bool valid = row_thread && row < R;
  const int env_l = valid ? (int)(row / sh.N) : 0;
  const int cand_l = valid ? (int)(row - (long long)env_l * sh.N) : 0;
  float cost = 0.0f;

  if (row_thread) {
    for (int o = 0; o < O; ++o) bufS[o * TM + t] = valid ? dep_load(s0 + (long long)env_l * O + o) : 0.0f;
  }

  for (int h = 0; h < sh.H; ++h) {
    float act_cost = 0.0f, a0 = 0.0f, ctl_sum = 0.0f;
    if (row_thread) {
      // normalize_state: (s - mean) / std   (data.py:258-260)
      for (int o = 0; o < O; ++o)
        bufA[o * TM + t] = __fdiv_rn(__fsub_rn(bufS[o * TM + t], __ldg(m.mu_s + o)), __ldg(m.sd_s + o));
      if (valid) {
        float* aout = actions_out ? actions_out + ((long long)h * R + row) * A : nullptr;
        for_each_action(src, A, sh.H, h, env_l, cand_l, row, R, [&](int a, float v) {
          const float vn = __fdiv_rn(__fsub_rn(v, __ldg(m.mu_a + a)), __ldg(m.sd_a + a));
          bufA[(O + a) * TM + t] = vn;
          bufN[a * TM + t] = vn;
          act_cost = __fadd_rn(act_cost, cosh_term(v, m.beta));
          if (a == 0) a0 = v;
          ctl_sum += fabsf(v) < 1.0f ? 1.0f - v * v : 0.0f;  // quadratic tolerance of the control (task costs)
          if (aout) aout[a] = v;
        });
      } else {
        for (int a = 0; a < A; ++a) { bufA[(O + a) * TM + t] = 0.0f; bufN[a * TM + t] = 0.0f; }
      }
    }
    __syncthreads();
    dense_layer<TM, CPT, CG, true>(m.W1t, m.b1, bufA, bufB, D, U);
    __syncthreads();
    dense_layer<TM, CPT, CG, true>(m.W2t, m.b2, bufB, bufA, U, U);
    __syncthreads();
    dense_layer<TM, 1, CG, false>(m.W3t, m.b3, bufA, bufB, U, O);
    __syncthreads();
    if (row_thread) {
      float st_cost = 0.0f;
      float* sout = (states_out && valid) ? states_out + ((long long)h * R + row) * O : nullptr;
      const bool smooth = m.cost_kind == MBRL_COST_SMOOTHABS_COSH;
      for (int o = 0; o < O; ++o) {
        // unnormalize_state: y * std + mean   (data.py:255-257)
        const float s = __fadd_rn(__fmul_rn(bufB[o * TM + t], __ldg(m.sd_s + o)), __ldg(m.mu_s + o));
        bufS[o * TM + t] = s;
        if (smooth) st_cost = __fadd_rn(st_cost, smooth_abs_term(s, __ldg(m.goal + o), __ldg(m.cost_w + o), m.alpha, m.alpha2));
        if (sout) sout[o] = s;
      }
      if (smooth) {
        // CoshLoss: beta^2 * mean_a(cosh(a/beta) - 1); row cost pairs s_{h+1} with a_h
        const float ac = __fmul_rn(m.beta2, __fdiv_rn(act_cost, (float)A));
        cost = __fadd_rn(cost, __fadd_rn(st_cost, ac));
      } else if (is_task_cost(m.cost_kind)) {  // the host checked that O covers the picked entries
        int pick[4];
        task_pick_indices(m.cost_kind, pick);
        float p4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) p4[q] = pick[q] >= 0 ? bufS[pick[q] * TM + t] : 0.0f;
        cost += task_cost(m.cost_kind, p4, a0, ctl_sum / (float)A);
      }
    }
    if (reward_head) {
      // RewardAgent (src/mbrl/agents.py:349-358): cost(s_{h+1}, a_h) = unnormalize_reward(linear4(trunk(
      // [normalize(s_{h+1}), normalize(a_h)]))) -- a second trunk evaluation (models.py:135-163).
      // Layer 3 (the last reader of bufA) is behind the barrier above.
      if (row_thread) {
        for (int o = 0; o < O; ++o)
          bufA[o * TM + t] = __fdiv_rn(__fsub_rn(bufS[o * TM + t], __ldg(m.mu_s + o)), __ldg(m.sd_s + o));
        for (int a = 0; a < A; ++a) bufA[(O + a) * TM + t] = bufN[a * TM + t];
      }
      __syncthreads();
      dense_layer<TM, CPT, CG, true>(m.W1t, m.b1, bufA, bufB, D, U);
      __syncthreads();
      dense_layer<TM, CPT, CG, true>(m.W2t, m.b2, bufB, bufA, U, U);
      __syncthreads();
      if (row_thread) {
        float r = 0.0f;
        for (int k = 0; k < U; ++k) r = fmaf(bufA[k * TM + t], __ldg(m.W4 + k), r);
        r = __fadd_rn(r, m.b4);
        cost = __fadd_rn(cost, __fadd_rn(__fmul_rn(r, m.sd_r), m.mu_r));  // unnormalize_field (data.py:255-257)
      }
      __syncthreads();  // bufA is rewritten by the row threads at the top of the next step
    }
    // bufA is rewritten by row threads next step: every warp has passed the barrier after
    // layer 3, which was the last reader of bufA.
  }
  if (valid) costs[row] = cost;
}

}  // namespace mbrl
