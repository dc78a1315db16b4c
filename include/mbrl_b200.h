/*
 * mbrl_b200.h -- C ABI of the B200-native MPC planning hot path.
 *
 * Drop-in boundary for the planning path of Khodeir/mujoco-mbrl.  The reference has no
 * native layer at all (pure Python/PyTorch-CPU), so there is no existing FFI to mirror;
 * each entry point below names the reference Python interface it replaces (paths are
 * relative to the reference root).  INTEGRATION.md shows the ctypes binding a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative MBRL_E_* code on failure;
 *     mbrl_last_error() returns a thread-local human-readable message;
 *   - plain pointers and sizes only; `h_` = host pointer, `d_` = device pointer
 *     (same CUDA primary context, e.g. torch tensor.data_ptr());
 *   - `stream` is a cudaStream_t passed as void* (NULL = the CUDA default stream); work is
 *     only enqueued, never synchronised, except in mbrl_plan which runs on the handle's own
 *     stream and returns after the plan has reached the host;
 *   - caller owns every buffer it passes; the handle owns device weights and scratch;
 *   - one handle per (device, shape); a handle is not thread-safe;
 *   - all fp32 matrices are row-major; nn.Linear weights are [out, in];
 *   - trajectories use the reference's flat step-major layout: row h*R + r is row r at
 *     step h (src/mbrl/planners.py:199-209) with R = num_envs * num_candidates and
 *     r = env * num_candidates + candidate;
 *   - there is NO CPU fallback: every entry point that computes requires a CUDA device
 *     and fails with MBRL_E_CUDA otherwise.
 */
#ifndef MBRL_B200_H_
#define MBRL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBRL_ABI_VERSION 2

/* error codes */
#define MBRL_OK 0
#define MBRL_E_INVALID (-1)   /* bad argument / unsupported shape */
#define MBRL_E_CUDA (-2)      /* CUDA runtime error (message has the cudaError string) */
#define MBRL_E_STATE (-3)     /* call order: weights / norm / cost not set yet */
#define MBRL_E_UNSUPPORTED (-4)

/* rollout engines */
#define MBRL_ENGINE_SIMT_FP32 0 /* fp32 CUDA-core kernel: 1e-5 parity cross-check      */
#define MBRL_ENGINE_TC_BF16 1   /* tcgen05 tensor cores, bf16 operands, fp32 accumulate */
#define MBRL_ENGINE_TC_FP16 2   /* tcgen05 tensor cores, fp16 operands, fp32 accumulate */

/* where a candidate's actions come from */
#define MBRL_SAMPLE_INJECT_ACTIONS 0 /* d_injected holds final actions [H*R, A]                */
#define MBRL_SAMPLE_INJECT_NOISE 1   /* d_injected holds N(0,1) draws; a = clip(mu + sd*z)      */
#define MBRL_SAMPLE_GAUSSIAN 2       /* Philox4x32-10 + Box-Muller; a = clip(mu + sd*z)         */
#define MBRL_SAMPLE_UNIFORM 3        /* Philox4x32-10; a = lo + (hi-lo)*u   (reference sampler
                                        distribution, src/mbrl/env_wrappers.py:50-62)          */

/* cost kinds (epilogue of the rollout kernel) */
#define MBRL_COST_SMOOTHABS_COSH 0 /* state_action_cost = SmoothAbsLoss + CoshLoss
                                      (src/mbrl/agents.py:182-183, models.py:244-272)          */
#define MBRL_COST_DMC_CARTPOLE_SWINGUP 1 /* 1 - smooth cartpole reward restated on the observation
                                      [x, cos, sin, x_dot, theta_dot] and the control
                                      (dm_control/suite/cartpole.py:216-226, utils/rewards.py:88-130);
                                      not used by the reference planner;
                                      every engine; weights/goal are ignored                   */
#define MBRL_COST_DMC_HUMANOID_RUN 3 /* 1 - Humanoid.get_reward at move_speed 10 restated on the egocentric
                                      observation (head_height obs[21], torso zz obs[36], com velocity
                                      obs[37:39]) and the control (dm_control/suite/humanoid.py:172-211);
                                      every engine; weights/goal are ignored                   */
#define MBRL_COST_DMC_CHEETAH_RUN 4 /* 1 - Cheetah.get_reward (dm_control/suite/cheetah.py:91-97): linear
                                      tolerance of the forward speed up to 10 m/s.  The reward's
                                      torso_subtreelinvel sensor is not part of the observation
                                      (cheetah.py:59-61,83-89): the root-x joint velocity obs[8] stands in
                                      for it (a documented PROXY); weights/goal ignored */
#define MBRL_COST_DMC_WALKER_WALK 5 /* 1 - PlanarWalker.get_reward at move_speed 1 (walker.py:135-158):
                                      torso height obs[14], torso zz obs[0] (planar: xx == zz), and the
                                      root-x joint velocity obs[16] as PROXY for the unobserved
                                      horizontal-velocity sensor; weights/goal ignored               */
#define MBRL_COST_REWARD_HEAD 2 /* RewardAgent's cost (src/mbrl/agents.py:342-366): a second trunk
                                      evaluation at (s_{h+1}, a_h) through ModelWithReward's
                                      linear4 head, un-normalised with the reward statistics
                                      (src/mbrl/models.py:125-163, data.py:255-257); minimised like
                                      a cost, as the reference does.  Needs mbrl_set_reward_head;
                                      weights/goal are ignored.  On the tensor-core engines the handle
                                      switches to the weight-streaming kernel (hidden <= 512), which
                                      runs the second trunk pass on the tensor cores              */

typedef struct MbrlPlanner MbrlPlanner;

typedef struct MbrlConfig {
  int32_t obs_dim;        /* O: state_dim of Model (src/mbrl/models.py:96)                     */
  int32_t act_dim;        /* A                                                                  */
  int32_t hidden;         /* U: hidden_units (src/mbrl/models.py:97)                            */
  int32_t horizon;        /* H: MPCPolicy.horizon (src/mbrl/agents.py:33)                       */
  int32_t num_candidates; /* N on this GPU: num_trajectories (src/mbrl/planners.py:141,153)    */
  int32_t num_envs;       /* E independent planning problems batched on this GPU (>= 1)        */
  int32_t max_iterations; /* I_max: CEM iterations a plan may use (1 = random shooting only)   */
  int32_t max_elites;     /* k_max                                                              */
  int32_t engine;         /* MBRL_ENGINE_*                                                      */
  int32_t device;         /* CUDA device ordinal                                                */
} MbrlConfig;

typedef struct MbrlPlanArgs {
  int32_t iterations;   /* 1 = random shooting (argmin); >1 = CEM                              */
  int32_t elites;       /* k (ties -> lower index); forced to 1 when iterations == 1           */
  int32_t sample_mode;  /* MBRL_SAMPLE_*                                                        */
  int32_t return_mean;  /* 0: return best-ever candidate; 1: return final mean sequence        */
  uint64_t seed;        /* Philox key                                                           */
  uint32_t cand_offset; /* global index of this shard's first candidate (population sharding)  */
  uint32_t env_offset;  /* global index of this shard's first environment (env sharding)       */
  const float* h_injected; /* [iterations, H*R, A] host (INJECT_* modes) or NULL               */
  const float* h_mu0;      /* [E, H, A] host initial mean, NULL -> (lo+hi)/2                    */
  const float* h_sd0;      /* [E, H, A] host initial std,  NULL -> (hi-lo)/2                    */
  int32_t actions_only; /* 1: emit only the action sequence; the fp32 replay that produces the
                           predicted states is skipped and out_states is zero-filled.  MPCPolicy
                           only consumes plan[1][0] (src/mbrl/agents.py:56)                      */
  int32_t warm_start;   /* MBRL_WARM_* bit mask: warm start across MPC steps with the sampling mean
                           RESIDENT ON THE DEVICE (MPCPolicy hands the previous plan to the next call,
                           src/mbrl/agents.py:41-47).  Ignored when h_mu0 is given. */
  float warm_std;       /* std of a warm-started distribution; <= 0 -> (hi-lo)/2                 */
  int32_t reserved;
} MbrlPlanArgs;

#define MBRL_WARM_USE 1  /* seed the mean with the previous plan's final mean shifted by one step (last
                            step repeated), if the handle holds one; std = warm_std                */
#define MBRL_WARM_KEEP 2 /* refit after the last iteration and keep this plan's final mean in the
                            handle for the next call (without this bit the stored mean is dropped:
                            episode start, agents.py:38-40)                                        */

typedef struct MbrlPlanInfo {
  float best_cost;
  int32_t best_iteration;
  int32_t best_index; /* candidate index within the environment (local to this shard)          */
  int32_t reserved;
} MbrlPlanInfo;

int mbrl_abi_version(void);
const char* mbrl_last_error(void);

/* Life cycle.  Replaces constructing Model(...) + MPCPolicy(...) state
 * (src/mbrl/models.py:96-104, src/mbrl/agents.py:29-36). */
int mbrl_create(const MbrlConfig* cfg, MbrlPlanner** out);
int mbrl_destroy(MbrlPlanner* p);

/* Model.linear{1,2,3}.{weight,bias} (src/mbrl/models.py:99-101); host fp32, [out,in].
 * Call again whenever the host retrains the model (src/mbrl/models.py:84-86). */
int mbrl_set_weights(MbrlPlanner* p, const float* h_W1, const float* h_b1, const float* h_W2,
                     const float* h_b2, const float* h_W3, const float* h_b3);
/* stats["observations"|"actions"]["mean"|"std"] used by normalize_field /
 * unnormalize_field (src/mbrl/data.py:255-269).  NULL pointers mean identity (0 / 1). */
int mbrl_set_norm(MbrlPlanner* p, const float* h_mu_s, const float* h_sd_s, const float* h_mu_a,
                  const float* h_sd_a);
/* SmoothAbsLoss(weights, goal_state, alpha) + CoshLoss(beta)
 * (src/mbrl/models.py:244-272; goal mutates via set_goal_state, models.py:240-241). */
int mbrl_set_cost(MbrlPlanner* p, int32_t kind, const float* h_weights, const float* h_goal,
                  double alpha, double beta); /* doubles: the reference squares the Python
                                                 floats in fp64 before the fp32 tensor op */
/* bounds EnvWrapper._sample_action derives from the action spec
 * (src/mbrl/env_wrappers.py:52-55: dimension 0's bounds for every dim, clipped to +-3). */
int mbrl_set_action_bounds(MbrlPlanner* p, float lo, float hi);
/* ModelWithReward.linear4 (src/mbrl/models.py:132: weight [1, U], bias [1]) and the "rewards"
 * entry of the dataset statistics that unnormalize_reward closes over (agents.py:347). */
int mbrl_set_reward_head(MbrlPlanner* p, const float* h_W4, float b4, float reward_mean, float reward_std);

/* One whole plan with HOST buffers: replaces RandomShootingPlanner.plan
 * (src/mbrl/planners.py:143-187) for iterations == 1 and adds CEM for iterations > 1.
 *   h_s0          [E, O]      initial_state
 *   h_out_states  [E, H, O]   predicted s_1..s_H of the chosen sequence (s_0 excluded,
 *                             as the reference returns them, planners.py:212-215)
 *   h_out_actions [E, H, A]
 *   h_info        [E]         nullable
 *   h_out_mu/sd   [E, H, A]   nullable: final sampling distribution (warm start)
 * H2D of s0 and D2H of the plan happen inside; returns after the plan is on the host. */
int mbrl_plan(MbrlPlanner* p, const MbrlPlanArgs* args, const float* h_s0, float* h_out_states,
              float* h_out_actions, MbrlPlanInfo* h_info, float* h_out_mu, float* h_out_sd);

/* Same plan with everything resident in HBM, enqueued on `stream` without host sync
 * (inputs/outputs are device pointers; d_injected replaces args->h_injected). */
int mbrl_plan_device(MbrlPlanner* p, const MbrlPlanArgs* args, const float* d_s0,
                     const float* d_injected, float* d_out_states, float* d_out_actions,
                     MbrlPlanInfo* d_info, void* stream);

/* Batched GradientDescentPlanner (src/mbrl/planners.py:28-137): Adam on `restarts` independent action
 * sequences, back-propagation through the H-step model rollout, the reference's early stop.  Uses the
 * handle's fp32 model / normalisers / SmoothAbs + Cosh cost (any engine; num_envs must be 1).
 *   h_s0            [O]               initial state, shared by all restarts
 *   h_init_actions  [restarts, H, A]  initial sequences (the reference draws one with sample_action,
 *                                     planners.py:88-100, or takes initial_trajectory's)
 *   h_out_states    [restarts, H+1, O]  s_0 first, then the states of the LAST forward pass -- computed with
 *                                     the actions before the final Adam step, as the reference returns them
 *   h_out_actions   [restarts, H, A]  the updated actions
 *   h_out_cost      [restarts]        nullable: loss of that last forward pass
 *   h_out_iters     [restarts]        nullable: iterations run (early stop)                          */
typedef struct MbrlGdArgs {
  int32_t restarts;
  int32_t iterations;   /* num_iterations (planners.py:29: 40)                     */
  float lr;             /* Adam learning rate (planners.py:114: 0.01)              */
  float stop_condition; /* mean |a_old - a_new| below which it stops (planners.py:29: 0.002) */
  float beta1, beta2, eps; /* torch.optim.Adam defaults: 0.9, 0.999, 1e-8         */
  int32_t reserved;
} MbrlGdArgs;
int mbrl_plan_gd(MbrlPlanner* p, const MbrlGdArgs* args, const float* h_s0, const float* h_init_actions,
                 float* h_out_states, float* h_out_actions, float* h_out_cost, int32_t* h_out_iters);

/* ---- building blocks on device buffers (parity tests, sharded host loops) ---- */

/* The hot loop of _generate_trajectories (src/mbrl/planners.py:199-210) fused with
 * DynamicsModel.forward (models.py:13-29) and the cost (agents.py:182-183):
 *   d_s0 [E,O]; d_injected [H*R,A] or NULL; d_mu/d_sd [E,H,A] (Gaussian / noise modes);
 *   d_costs [R] out; d_states_out [H*R,O] / d_actions_out [H*R,A] nullable (debug/parity:
 *   production plans never materialise trajectories). */
int mbrl_rollout(MbrlPlanner* p, int32_t sample_mode, uint64_t seed, uint32_t iteration,
                 uint32_t cand_offset, uint32_t env_offset, const float* d_s0,
                 const float* d_injected, const float* d_mu, const float* d_sd, float* d_costs,
                 float* d_states_out, float* d_actions_out, void* stream);

/* Materialise the sampler's output: d_out [H*R, A] (Gaussian / uniform Philox modes). */
int mbrl_sample(MbrlPlanner* p, int32_t sample_mode, uint64_t seed, uint32_t iteration,
                uint32_t cand_offset, uint32_t env_offset, const float* d_mu, const float* d_sd,
                float* d_out, void* stream);
/* Raw Philox4x32-10 words for known-answer tests: d_ctr [n,4], d_key [n,2] -> d_out [n,4]. */
int mbrl_philox_raw(const uint32_t* d_ctr, const uint32_t* d_key, uint32_t* d_out, int64_t n,
                    void* stream);

/* Segmented elite select: for each of `segments` cost arrays of length n pick the k
 * smallest, ties toward the lower index (consistent with np.argmin, planners.py:184).
 *   d_elite_idx  [segments, k] int32, ascending index order
 *   d_elite_cost [segments, k] nullable
 *   d_best       [segments] nullable: (cost, -, index) of the minimum                  */
int mbrl_topk(const float* d_costs, int32_t segments, int32_t n, int32_t k, int32_t* d_elite_idx,
              float* d_elite_cost, MbrlPlanInfo* d_best, void* stream);

/* Mean / population-std refit of the sampling distribution from the elite candidates'
 * action sequences (regenerated from the sampler, or gathered from d_injected):
 *   d_mu_new/d_sd_new [E,H,A]. */
int mbrl_refit(MbrlPlanner* p, int32_t sample_mode, uint64_t seed, uint32_t iteration,
               uint32_t cand_offset, uint32_t env_offset, const float* d_injected,
               const float* d_mu, const float* d_sd, const int32_t* d_elite_idx, int32_t k,
               float* d_mu_new, float* d_sd_new, void* stream);

/* Emit chosen plans: regenerate the action sequence of candidate d_best[e].best_index drawn
 * in iteration d_best[e].best_iteration (from d_mu_hist/d_sd_hist [iterations+1, E, H, A],
 * slot i = distribution sampled in iteration i) -- or the final mean when return_mean -- and
 * replay it through the fp32 model: d_out_states [E,H,O], d_out_actions [E,H,A].  This is
 * the (states, actions) pair RandomShootingPlanner.plan returns (planners.py:184-187). */
int mbrl_emit(MbrlPlanner* p, int32_t sample_mode, uint64_t seed, uint32_t cand_offset,
              uint32_t env_offset, const float* d_s0, const float* d_injected,
              const float* d_mu_hist, const float* d_sd_hist, int32_t iterations,
              int32_t return_mean, const MbrlPlanInfo* d_best, float* d_out_states,
              float* d_out_actions, void* stream);

/* ---- population sharding over the GPUs of one node (one process per GPU, NCCL) ----------
 * After mbrl_comm_init the handle is one shard of a population of world * num_candidates
 * candidates (rank r owns global candidates [r*N, (r+1)*N)); mbrl_plan / mbrl_plan_device then
 * run the sharded CEM loop entirely on the stream: per iteration one ncclAllGather of each
 * rank's min(k, N) cheapest (cost, global index) pairs, the same global top-k selection on every
 * rank (ties -> lower global index), and a redundant refit that regenerates the elites from
 * their global indices (Philox counters carry the global index), so no second collective is
 * needed and every rank holds bit-identical mean/std and emits the same plan.  args->elites is
 * the GLOBAL k; info.best_index is a GLOBAL candidate index; num_envs must be 1; Philox sample
 * modes only.  Ranks send their expected share of the elites plus 8 sigma (the shards are i.i.d.) and the
 * merge verifies on the device that this was exact; otherwise info.reserved = 1 and mbrl_plan
 * transparently redoes the plan with worst-case-size gathers (mbrl_plan_device leaves the flag
 * to the caller).  libnccl.so.2 is resolved at run time (dlopen), not at link time.
 *   mbrl_nccl_unique_id: rank 0 fills 128 bytes, the host broadcasts them to all ranks.        */
/* Peer-memory transport for the same sharded loop (NVLink P2P through CUDA IPC instead of the
 * ncclAllGather): every rank exports one gather buffer (mbrl_p2p_export: 64-byte IPC handle),
 * the host all-gathers the handles, mbrl_p2p_attach opens the peers' buffers.  Everything in a
 * buffer is a packet -- a 64-bit {value, sequence tag} word written with one scalar store -- so
 * data is consumed as it lands, without flags or system-scope fences.  Per iteration three
 * kernels run: the local select stores this rank's cheapest costs straight into every peer's
 * buffer; the merge select (resident and polling while the rollout still runs) stages the
 * gathered costs, finds the global threshold and keeps this rank's own elites; the refit sums
 * them, exchanges the H*ceil(A/4)*8 partial sums the same way and adds the ranks' partials in
 * rank order.  Buffers are double-buffered by iteration parity; a writer can never be two
 * iterations ahead of a reader because its own merge needs every rank's packets of the iteration
 * in between.  Every wait gives up after MBRL_P2P_TIMEOUT_S (default 120) seconds of wall clock:
 * the plan then fails instead of using stale data.  Falls back to NCCL when no peer buffers are
 * attached.                                                                                     */
int mbrl_p2p_export(MbrlPlanner* p, int32_t world, uint8_t* h_handle64);
int mbrl_p2p_attach(MbrlPlanner* p, const uint8_t* h_handles, int32_t rank, int32_t world);
/* Closes whatever peer buffers were opened (also after a failed attach) so that the handle can
 * take the NCCL transport instead; the exported buffer stays allocated.  Idempotent. */
int mbrl_p2p_detach(MbrlPlanner* p);
/* Summation order of the refit.  A plan sharded over W ranks adds W per-rank partial sums
 * (rank r: the elites among candidates [r*N, (r+1)*N)) in rank order; an UNSHARDED planner over
 * the W*N candidates reproduces it bit for bit after mbrl_set_refit_segments(p, W) (W must divide
 * its population; single-environment planners, W <= 64).  1 (the default) = the plain order.
 * Test / verification knob: the result is the same mean and std up to fp32 rounding.           */
int mbrl_set_refit_segments(MbrlPlanner* p, int32_t segments);
int mbrl_nccl_unique_id(uint8_t* h_id128);
int mbrl_comm_init(MbrlPlanner* p, const uint8_t* h_id128, int32_t rank, int32_t world);
int mbrl_comm_destroy(MbrlPlanner* p);

/* Diagnostic for the tensor-core engines (tests only): enable != 0 arms a dump of the raw
 * fp32 accumulators of row tile 0 at step 0 ([3 layers][128 rows][512 cols] floats) followed
 * by a clock64() timeline of tile 1 ([64 steps][32 events] int64) by the next mbrl_rollout;
 * h_out != NULL copies the dump (3*128*512*4 + 64*32*8 bytes) to the host after a sync. */
int mbrl_tc_debug(MbrlPlanner* p, int32_t enable, float* h_out);

#ifdef __cplusplus
}
#endif
#endif /* MBRL_B200_H_ */
