// Shared device-side definitions for the planning kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mbrl_b200.h"

namespace mbrl {

constexpr int kMaxObs = 128;
constexpr int kMaxAct = 32;
constexpr int kMaxHidden = 1024;

// Device view of the dynamics model, normalisers and cost
// (DynamicsModel.forward src/mbrl/models.py:13-29; data.py:255-260; models.py:244-272).
struct ModelDev {
  int O, A, D, U;
  const float* W1t;  // [D][U]  transposed nn.Linear weights (K-major: coalesced over outputs)
  const float* b1;   // [U]
  const float* W2t;  // [U][U]
  const float* b2;   // [U]
  const float* W3t;  // [U][O]
  const float* b3;   // [O]
  const float* mu_s;  // [O]
  const float* sd_s;  // [O]
  const float* mu_a;  // [A]
  const float* sd_a;  // [A]
  const float* cost_w;  // [O]
  const float* goal;    // [O]
  float alpha, alpha2;  // alpha2 = (float)(alpha_fp64^2), as torch folds the Python scalar
  float beta, beta2;
  int cost_kind;
  const float* W4;  // [U] reward head (MBRL_COST_REWARD_HEAD), else null
  float b4, mu_r, sd_r;
};

// Problem extents on this GPU: R = E*N rows, row r = env*N + cand.
struct Shape {
  int H, N, E;
  __host__ __device__ long long rows() const { return (long long)N * E; }
};

// Where actions come from (MBRL_SAMPLE_*).
struct ActionSource {
  int mode;
  const float* buf;  // injected actions or N(0,1) noise, [H*R, A] step-major
  const float* mu;   // [E, H, A]
  const float* sd;   // [E, H, A]
  uint32_t seed_lo, seed_hi;
  uint32_t iteration;
  uint32_t cand_offset;  // global index of local candidate 0
  uint32_t env_offset;   // global index of local env 0
  float lo, hi;
};

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor is still running; pdl_wait() blocks until every
// prerequisite grid has completed and its writes are visible, pdl_trigger() lets the successor
// start its own prologue.  Both are no-ops for ordinary launches.
// IMPORTANT: data produced by a predecessor kernel must be read with dep_load() (ld.global.cg), never
// with __ldg / through a const __restrict__ pointer: ptxas treats non-coherent loads (LDG.CONSTANT)
// as freely movable and hoists them ABOVE griddepcontrol.wait (seen in SASS: the refit kernel read
// the old mean before its producer had written it), and the L1 line may be stale besides.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename T>
__device__ __forceinline__ T dep_load(const T* p) { return __ldcg(p); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ float clipf(float x, float lo, float hi) {
  return fminf(fmaxf(x, lo), hi);
}

// SmoothAbsLoss term for one state dim: sqrt(((s-g)*w)^2 + alpha^2) - alpha
// (src/mbrl/models.py:255-259), separate roundings like the torch op chain.
__device__ __forceinline__ float smooth_abs_term(float s, float g, float w, float alpha,
                                                 float alpha2) {
  float x = __fmul_rn(__fsub_rn(s, g), w);
  return __fsub_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(x, x), alpha2)), alpha);
}

// CoshLoss term for one action dim: cosh(a/beta) - 1   (src/mbrl/models.py:271-272)
__device__ __forceinline__ float cosh_term(float a, float beta) {
  return __fsub_rn(coshf(__fdiv_rn(a, beta)), 1.0f);
}

// 1 - cartpole swing-up reward (dm_control/suite/cartpole.py:216-226) from x, cos(theta),
// theta_dot and the control; rewards.tolerance with the default gaussian sigmoid is
// exp(ln(0.1) * (d/margin)^2), the quadratic one with value_at_margin 0 is 1 - d^2 inside |d| < 1.
__device__ __forceinline__ float dmc_cartpole_cost(float x, float cosang, float thdot, float a0) {
  const float ln01 = -2.302585092994046f;
  const float upright = 0.5f * (cosang + 1.0f);
  const float xc = 0.5f * x, vc = 0.2f * thdot;
  const float centered = 0.5f * (1.0f + expf(ln01 * xc * xc));
  const float small_velocity = 0.5f * (1.0f + expf(ln01 * vc * vc));
  const float ctl = fabsf(a0) < 1.0f ? 1.0f - a0 * a0 : 0.0f;
  const float small_control = 0.2f * (4.0f + ctl);
  return 1.0f - upright * small_control * small_velocity * centered;
}

// 1 - Humanoid.get_reward, move_speed = 10 (dm_control/suite/humanoid.py:187-211): standing =
// tolerance(head_height, [1.4, inf), margin 0.35, gaussian); upright = tolerance(torso zz, [0.9, inf),
// margin 1.9, linear, value_at_margin 0); small_control = (4 + mean_a(1 - a^2 inside |a| < 1)) / 5;
// move = (5 * tolerance(|com_vel_xy|, [10, inf), margin 10, linear, 0) + 1) / 6.
__device__ __forceinline__ float dmc_humanoid_run_cost(float head, float zz, float vx, float vy, float ctl_mean) {
  const float ln01 = -2.302585092994046f;
  const float dh = (1.4f - head) * (1.0f / 0.35f);
  const float standing = head >= 1.4f ? 1.0f : expf(ln01 * dh * dh);
  const float du = (0.9f - zz) * (1.0f / 1.9f);
  const float upright = zz >= 0.9f ? 1.0f : (du < 1.0f ? 1.0f - du : 0.0f);
  const float small_control = 0.2f * (4.0f + ctl_mean);
  const float speed = sqrtf(vx * vx + vy * vy);
  const float dm = (10.0f - speed) * 0.1f;
  const float move_raw = speed >= 10.0f ? 1.0f : (dm < 1.0f ? 1.0f - dm : 0.0f);
  const float move = (5.0f * move_raw + 1.0f) * (1.0f / 6.0f);
  return 1.0f - small_control * standing * upright * move;
}

// 1 - Cheetah.get_reward (dm_control/suite/cheetah.py:91-97): tolerance(speed, [10, inf), margin 10,
// linear, value_at_margin 0).  physics.speed() is the torso_subtreelinvel sensor, which is NOT in the
// observation (cheetah.py:59-61): the documented proxy is the root-x joint velocity obs[8]
// (obs = qpos[1:] | qvel, cheetah.py:83-89).
__device__ __forceinline__ float dmc_cheetah_run_cost(float speed) {
  const float d = (10.0f - speed) * 0.1f;
  const float r = speed >= 10.0f ? 1.0f : (d < 1.0f ? 1.0f - d : 0.0f);
  return 1.0f - r;
}

// 1 - PlanarWalker.get_reward at move_speed 1 (walker-walk; dm_control/suite/walker.py:135-158):
// standing = tolerance(torso_height, [1.2, inf), margin 0.6, gaussian); upright = (1 + torso zz) / 2;
// stand = (3 standing + upright) / 4; move = tolerance(v, [1, inf), margin 0.5, linear, value_at_margin
// 0.5); reward = stand * (5 move + 1) / 6.  torso_height = obs[14], zz = obs[0] (planar: xx == zz);
// the horizontal-velocity sensor is not observed: proxy = root-x joint velocity obs[16]
// (walker.py:88-102, qvel order rootz, rootx, rooty).
__device__ __forceinline__ float dmc_walker_walk_cost(float height, float zz, float v) {
  const float ln01 = -2.302585092994046f;
  const float dh = (1.2f - height) * (1.0f / 0.6f);
  const float standing = height >= 1.2f ? 1.0f : expf(ln01 * dh * dh);
  const float upright = 0.5f * (1.0f + zz);
  const float stand = 0.25f * (3.0f * standing + upright);
  const float sx = (1.0f - v);  // d / margin * (1 - value_at_margin) = (1 - v) / 0.5 * 0.5
  const float move = v >= 1.0f ? 1.0f : (sx < 1.0f ? 1.0f - sx : 0.0f);
  return 1.0f - stand * (5.0f * move + 1.0f) * (1.0f / 6.0f);
}

// The dm_control task costs read at most four entries of the un-normalised predicted state: the
// tensor-core engines pick them out of the output epilogue instead of keeping the whole state.
__host__ __device__ inline void task_pick_indices(int kind, int (&idx)[4]) {
  idx[0] = idx[1] = idx[2] = idx[3] = -1;
  if (kind == MBRL_COST_DMC_CARTPOLE_SWINGUP) { idx[0] = 0; idx[1] = 1; idx[2] = 4; }
  else if (kind == MBRL_COST_DMC_HUMANOID_RUN) { idx[0] = 21; idx[1] = 36; idx[2] = 37; idx[3] = 38; }
  else if (kind == MBRL_COST_DMC_CHEETAH_RUN) { idx[0] = 8; }
  else if (kind == MBRL_COST_DMC_WALKER_WALK) { idx[0] = 14; idx[1] = 0; idx[2] = 16; }
}
// a0: the step's first control; ctl_mean: mean_a(1 - a^2 inside |a| < 1) of the step's controls.
__device__ __forceinline__ float task_cost(int kind, const float (&p)[4], float a0, float ctl_mean) {
  if (kind == MBRL_COST_DMC_CARTPOLE_SWINGUP) return dmc_cartpole_cost(p[0], p[1], p[2], a0);
  if (kind == MBRL_COST_DMC_HUMANOID_RUN) return dmc_humanoid_run_cost(p[0], p[1], p[2], p[3], ctl_mean);
  if (kind == MBRL_COST_DMC_CHEETAH_RUN) return dmc_cheetah_run_cost(p[0]);
  return dmc_walker_walk_cost(p[0], p[1], p[2]);
}
__host__ __device__ inline bool is_task_cost(int kind) {
  return kind == MBRL_COST_DMC_CARTPOLE_SWINGUP || kind == MBRL_COST_DMC_HUMANOID_RUN ||
         kind == MBRL_COST_DMC_CHEETAH_RUN || kind == MBRL_COST_DMC_WALKER_WALK;
}

}  // namespace mbrl
