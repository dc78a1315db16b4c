// C-ABI implementation (see include/mbrl_b200.h).  Host-side orchestration only: every
// byte of planning arithmetic happens in the kernels included below.  There is no CPU
// fallback: without a CUDA device every computing entry point returns MBRL_E_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "gd.cuh"
#include "philox.cuh"
#include "replay.cuh"
#include "rollout_simt.cuh"
#include "rollout_tc.cuh"
#include "rollout_tcf.cuh"
#include "select.cuh"

using namespace mbrl;

// --------------------------------------------------------------------------------------
// errors
// --------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define MBRL_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      return fail(MBRL_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));      \
    }                                                                                     \
  } while (0)
#define MBRL_REQUIRE(cond, msg) \
  do {                          \
    if (!(cond)) return fail(MBRL_E_INVALID, std::string("invalid argument: ") + (msg)); \
  } while (0)

// --------------------------------------------------------------------------------------
// NCCL, resolved at run time (the library must load on machines without NCCL / without a GPU)
// --------------------------------------------------------------------------------------
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
static NcclApi g_nccl;
static int load_nccl() {
  if (g_nccl.ok) return MBRL_OK;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // torch has usually loaded it already
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return fail(MBRL_E_UNSUPPORTED, std::string("cannot load libnccl.so.2: ") + dlerror());
  g_nccl.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(lib, "ncclCommInitRank");
  g_nccl.CommDestroy = (int (*)(NcclComm))dlsym(lib, "ncclCommDestroy");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(lib, "ncclAllGather");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather || !g_nccl.GetErrorString)
    return fail(MBRL_E_UNSUPPORTED, "libnccl.so.2 lacks a required symbol");
  g_nccl.ok = true;
  return MBRL_OK;
}
#define MBRL_NCCL(expr)                                                                         \
  do {                                                                                          \
    int r__ = (expr);                                                                           \
    if (r__ != 0) return fail(MBRL_E_CUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(r__)); \
  } while (0)
constexpr int kNcclUint32 = 3;  // ncclUint32 (nccl.h ncclDataType_t)

// --------------------------------------------------------------------------------------
// handle
// --------------------------------------------------------------------------------------
struct MbrlPlanner {
  MbrlConfig cfg{};
  int O = 0, A = 0, D = 0, U = 0, H = 0, N = 0, E = 0;
  long long R = 0;
  // device model (fp32, K-major transposed weights)
  float *W1t = nullptr, *b1 = nullptr, *W2t = nullptr, *b2 = nullptr, *W3t = nullptr, *b3 = nullptr;
  float *mu_s = nullptr, *sd_s = nullptr, *mu_a = nullptr, *sd_a = nullptr, *cost_w = nullptr, *goal = nullptr;
  float alpha = 0.4f, alpha2 = 0.16f, beta = 0.25f, beta2 = 0.0625f, lo = -1.f, hi = 1.f;
  int cost_kind = MBRL_COST_SMOOTHABS_COSH;
  bool have_weights = false, have_cost = false;
  // tensor-core engine state (packed 16-bit operand images), owned by rollout_tc.cuh
  TcModel tc{};
  std::vector<float> hW1, hb1, hW2, hb2, hW3, hb3;  // host copies: the tensor-core operand image is re-packed when the kernel variant changes
  // scratch
  float* d_s0 = nullptr;        // [E,O]
  float* d_costs = nullptr;     // [R]
  float* d_mu_hist = nullptr;   // [(Imax+1), E, H, A]
  float* d_sd_hist = nullptr;
  float* d_mu_last = nullptr;   // [E,H,A] final mean of the previous plan (MBRL_WARM_KEEP), resident for the next call's warm start
  bool have_last = false;
  int* d_elite = nullptr;       // [E, kmax]
  BestEver* d_best_ever = nullptr;  // [E]
  float* d_out_states = nullptr;    // [E,H,O]
  float* d_out_actions = nullptr;   // [E,H,A]
  MbrlPlanInfo* d_info = nullptr;   // [E]
  float* d_injected = nullptr;
  size_t injected_cap = 0;
  // pinned host staging
  float *h_s0 = nullptr, *h_out_states = nullptr, *h_out_actions = nullptr, *h_mu = nullptr, *h_sd = nullptr;
  MbrlPlanInfo* h_info = nullptr;
  cudaStream_t stream = nullptr;
  int num_sms = 0;
  size_t max_smem = 0;
  // population sharding (mbrl_comm_init)
  NcclComm comm = nullptr;
  int rank = 0, world = 1;
  float* d_ecost = nullptr;     // [k_l] local elite costs
  uint32_t* d_send = nullptr;   // [2*k_l]
  uint32_t* d_recv = nullptr;   // [world*2*k_l]
  float* d_gcost = nullptr;     // [world*k_l]
  int* d_gidx = nullptr;        // [world*k_l]
  int* d_pos = nullptr;         // [kmax]
  MbrlPlanInfo* d_best_now = nullptr;
  float* W4 = nullptr;  // reward head (mbrl_set_reward_head)
  float b4 = 0.f, mu_r = 0.f, sd_r = 1.f;
  bool have_reward_head = false;
  int* d_trunc = nullptr;       // truncation flag of the reduced-size elite gather
  float* d_refit_part = nullptr;           // [E][H*G][chunks][8] partial sums of the chunked refit
  unsigned int* d_refit_arrive = nullptr;  // [E][H*G] arrival counters (self-resetting)
  int refit_parts_cap = 0;                 // partial sums per slot that d_refit_part holds
  int refit_segments = 1;                  // mbrl_set_refit_segments: canonical summation order of the refit
  int* d_own_count = nullptr;              // population sharding: number of this rank's elites
  long long* d_shard_stamps = nullptr;     // MBRL_SHARD_TIMELINE diagnostic: [max_iterations][16] globaltimer stamps
  long long* d_plan_stamps = nullptr;      // MBRL_PLAN_TIMELINE diagnostic: [max_iterations][kPlanStampStride]
  long long* cur_stamps = nullptr;         // the current iteration's slice of d_plan_stamps (null: off)
  StageS0 stage_s0{nullptr, nullptr, 0};   // mbrl_plan: initial states staged by the plan's first kernel
  bool full_gather = false;     // force worst-case-size gathers (while a flagged plan is redone)
  int scratch_world = 0;        // world size the sharding scratch buffers were allocated for (0 = none)
  // peer-memory transport (mbrl_p2p_export / mbrl_p2p_attach)
  uint32_t* d_p2p_local = nullptr;  // exported packet buffer (layout: select.cuh, p2p_*_off)
  int p2p_slot = 0, p2p_world = 0;
  bool p2p_attached = false;
  P2pPeers p2p_peers{};
  int* d_p2p_error = nullptr;
  uint32_t p2p_seq = 0;
};

static ModelDev model_view(const MbrlPlanner* p) {
  ModelDev m;
  m.O = p->O; m.A = p->A; m.D = p->D; m.U = p->U;
  m.W1t = p->W1t; m.b1 = p->b1; m.W2t = p->W2t; m.b2 = p->b2; m.W3t = p->W3t; m.b3 = p->b3;
  m.mu_s = p->mu_s; m.sd_s = p->sd_s; m.mu_a = p->mu_a; m.sd_a = p->sd_a;
  m.cost_w = p->cost_w; m.goal = p->goal;
  m.alpha = p->alpha; m.alpha2 = p->alpha2; m.beta = p->beta; m.beta2 = p->beta2;
  m.cost_kind = p->cost_kind;
  m.W4 = p->W4; m.b4 = p->b4; m.mu_r = p->mu_r; m.sd_r = p->sd_r;
  return m;
}

static ActionSource action_source(const MbrlPlanner* p, int mode, uint64_t seed, uint32_t iteration,
                                  uint32_t cand_offset, uint32_t env_offset, const float* buf,
                                  const float* mu, const float* sd) {
  ActionSource s;
  s.mode = mode; s.buf = buf; s.mu = mu; s.sd = sd;
  s.seed_lo = (uint32_t)(seed & 0xFFFFFFFFull); s.seed_hi = (uint32_t)(seed >> 32);
  s.iteration = iteration; s.cand_offset = cand_offset; s.env_offset = env_offset;
  s.lo = p->lo; s.hi = p->hi;
  return s;
}

template <class T>
static cudaError_t dev_alloc(T** ptr, size_t count) {
  return cudaMalloc((void**)ptr, sizeof(T) * (count ? count : 1));
}

extern "C" int mbrl_abi_version(void) { return MBRL_ABI_VERSION; }
extern "C" const char* mbrl_last_error(void) { return g_err.c_str(); }

// MBRL_SHARD_TIMELINE=<prefix> (diagnostic): the sharded kernels of the last plan left globaltimer
// stamps per iteration; write them, in ns relative to the first, to <prefix>.rank<r>.txt:
//   select: start, packets sent | merge: start, all packets staged, end | refit (CTA 0): start, partial sums sent, all ranks' sums in, end
constexpr int kPlanStampCtas = 256;
constexpr int kPlanStampStride = 8 + 3 * kPlanStampCtas;
// MBRL_PLAN_TIMELINE=<path> (diagnostic, unsharded fused tcgen05 engine, <= 256 row tiles): globaltimer
// stamps of the last plan's kernels per iteration, in ns relative to the first:
//   top-k start (inputs ready), end | refit (CTA 0) start, end | rollout CTAs: entry, upstream data ready, exit
static void dump_plan_timeline(MbrlPlanner* p) {
  const char* path = getenv("MBRL_PLAN_TIMELINE");
  if (!path || !p->d_plan_stamps) return;
  const int I = p->cfg.max_iterations;
  std::vector<long long> h((size_t)kPlanStampStride * I);
  if (cudaDeviceSynchronize() != cudaSuccess ||
      cudaMemcpy(h.data(), p->d_plan_stamps, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return;
  FILE* f = std::fopen(path, "w");
  if (!f) return;
  long long t0 = 0;
  for (long long v : h) if (v && (!t0 || v < t0)) t0 = v;
  std::fprintf(f, "# ns since the first stamp; rollout columns: min / median / max over the CTAs\n");
  std::fprintf(f, "# it | rollout entry | rollout data-ready | rollout exit | topk start end | refit start end\n");
  for (int it = 0; it < I; ++it) {
    const long long* s = h.data() + (size_t)kPlanStampStride * it;
    if (!s[0]) break;
    std::fprintf(f, "%d |", it);
    for (int j = 0; j < 3; ++j) {
      std::vector<long long> v;
      for (int c = 0; c < kPlanStampCtas; ++c) if (s[8 + 3 * c + j]) v.push_back(s[8 + 3 * c + j] - t0);
      std::sort(v.begin(), v.end());
      if (v.empty()) std::fprintf(f, " - - - |");
      else std::fprintf(f, " %lld %lld %lld |", v.front(), v[v.size() / 2], v.back());
    }
    std::fprintf(f, " %lld %lld | %lld %lld\n", s[0] - t0, s[1] ? s[1] - t0 : -1, s[2] ? s[2] - t0 : -1, s[3] ? s[3] - t0 : -1);
  }
  std::fclose(f);
}

static void dump_shard_timeline(MbrlPlanner* p) {
  const char* prefix = getenv("MBRL_SHARD_TIMELINE");
  if (!prefix || !p->d_shard_stamps) return;
  const int I = p->cfg.max_iterations;
  std::vector<long long> h((size_t)16 * I);
  if (cudaDeviceSynchronize() != cudaSuccess ||
      cudaMemcpy(h.data(), p->d_shard_stamps, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return;
  const std::string path = std::string(prefix) + ".rank" + std::to_string(p->rank) + ".txt";
  if (FILE* f = std::fopen(path.c_str(), "w")) {
    std::fprintf(f, "# iteration: sel_start sel_pub | mrg_start mrg_acq mrg_end | rft_start rft_pub rft_acq rft_end   (ns since iteration 0 sel_start; absolute t0 %lld)\n", h[0]);
    for (int it = 0; it < I; ++it) {
      if (!h[(size_t)16 * it]) break;
      std::fprintf(f, "%d:", it);
      for (int j = 0; j < 9; ++j) std::fprintf(f, " %lld%s", h[(size_t)16 * it + j] ? h[(size_t)16 * it + j] - h[0] : -1, (j == 1 || j == 4) ? " |" : "");
      std::fprintf(f, "\n");
    }
    std::fclose(f);
  }
}

extern "C" int mbrl_destroy(MbrlPlanner* p) {
  if (!p) return MBRL_OK;
  cudaSetDevice(p->cfg.device);
  dump_shard_timeline(p);
  dump_plan_timeline(p);
  if (p->d_shard_stamps) cudaFree(p->d_shard_stamps);
  if (p->d_plan_stamps) cudaFree(p->d_plan_stamps);
  float* dev[] = {p->W1t, p->b1, p->W2t, p->b2, p->W3t, p->b3, p->mu_s, p->sd_s, p->mu_a, p->sd_a,
                  p->cost_w, p->goal, p->d_s0, p->d_costs, p->d_mu_hist, p->d_sd_hist, p->d_mu_last,
                  p->d_out_states, p->d_out_actions, p->d_injected};
  for (float* q : dev) if (q) cudaFree(q);
  if (p->d_elite) cudaFree(p->d_elite);
  if (p->W4) cudaFree(p->W4);
  if (p->d_refit_part) cudaFree(p->d_refit_part);
  if (p->d_refit_arrive) cudaFree(p->d_refit_arrive);
  if (p->d_own_count) cudaFree(p->d_own_count);
  if (p->d_best_ever) cudaFree(p->d_best_ever);
  if (p->d_info) cudaFree(p->d_info);
  if (p->comm && g_nccl.ok) g_nccl.CommDestroy(p->comm);
  if (p->p2p_attached)
    for (int r = 0; r < p->world; ++r)
      if (r != p->rank && p->p2p_peers.base[r]) cudaIpcCloseMemHandle(p->p2p_peers.base[r]);
  if (p->d_p2p_local) cudaFree(p->d_p2p_local);
  if (p->d_p2p_error) cudaFree(p->d_p2p_error);
  void* shard[] = {p->d_ecost, p->d_send, p->d_recv, p->d_gcost, p->d_gidx, p->d_pos, p->d_best_now, p->d_trunc};
  for (void* q : shard) if (q) cudaFree(q);
  tc_free(&p->tc);
  float* pinned[] = {p->h_s0, p->h_out_states, p->h_out_actions, p->h_mu, p->h_sd};
  for (float* q : pinned) if (q) cudaFreeHost(q);
  if (p->h_info) cudaFreeHost(p->h_info);
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
  return MBRL_OK;
}

extern "C" int mbrl_create(const MbrlConfig* cfg, MbrlPlanner** out) {
  if (!cfg || !out) return fail(MBRL_E_INVALID, "mbrl_create: null argument");
  *out = nullptr;
  MBRL_REQUIRE(cfg->obs_dim >= 1 && cfg->obs_dim <= kMaxObs, "obs_dim out of range [1,128]");
  MBRL_REQUIRE(cfg->act_dim >= 1 && cfg->act_dim <= kMaxAct, "act_dim out of range [1,32]");
  MBRL_REQUIRE(cfg->hidden >= 1 && cfg->hidden <= kMaxHidden, "hidden out of range [1,1024]");
  MBRL_REQUIRE(cfg->horizon >= 1 && cfg->horizon <= 4096, "horizon out of range");
  MBRL_REQUIRE(cfg->num_candidates >= 1, "num_candidates must be >= 1");
  MBRL_REQUIRE(cfg->num_envs >= 1, "num_envs must be >= 1");
  MBRL_REQUIRE((long long)cfg->num_envs * cfg->num_candidates < (1ll << 31), "E*N too large");
  MBRL_REQUIRE(cfg->max_iterations >= 1 && cfg->max_iterations <= 1024, "max_iterations out of range");
  MBRL_REQUIRE(cfg->max_elites >= 1 && (long long)cfg->max_elites <= 64ll * cfg->num_candidates,
               "max_elites out of range [1, N] (or [1, world*N] for a population shard)");
  MBRL_REQUIRE(cfg->engine >= MBRL_ENGINE_SIMT_FP32 && cfg->engine <= MBRL_ENGINE_TC_FP16, "unknown engine");
  int ndev = 0;
  MBRL_CUDA(cudaGetDeviceCount(&ndev));
  MBRL_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "no such CUDA device");
  MBRL_CUDA(cudaSetDevice(cfg->device));

  MbrlPlanner* p = new (std::nothrow) MbrlPlanner();
  if (!p) return fail(MBRL_E_INVALID, "out of host memory");
  p->cfg = *cfg;
  p->O = cfg->obs_dim; p->A = cfg->act_dim; p->D = p->O + p->A; p->U = cfg->hidden;
  p->H = cfg->horizon; p->N = cfg->num_candidates; p->E = cfg->num_envs;
  p->R = (long long)p->N * p->E;
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, cfg->device);
  if (e != cudaSuccess) { delete p; return fail(MBRL_E_CUDA, cudaGetErrorString(e)); }
  p->num_sms = prop.multiProcessorCount;
  p->max_smem = prop.sharedMemPerBlockOptin;
  if (cfg->engine != MBRL_ENGINE_SIMT_FP32 && prop.major != 10) {
    delete p;
    return fail(MBRL_E_UNSUPPORTED, "tcgen05 engines need an sm_100 device");
  }

  const int O = p->O, A = p->A, D = p->D, U = p->U, H = p->H, E = p->E;
  const size_t EHA = (size_t)E * H * A;
  bool ok = true;
  auto A_ = [&](cudaError_t err) { if (err != cudaSuccess) { ok = false; g_err = cudaGetErrorString(err); } };
  A_(dev_alloc(&p->W1t, (size_t)D * U)); A_(dev_alloc(&p->b1, U));
  A_(dev_alloc(&p->W2t, (size_t)U * U)); A_(dev_alloc(&p->b2, U));
  A_(dev_alloc(&p->W3t, (size_t)U * O)); A_(dev_alloc(&p->b3, O));
  A_(dev_alloc(&p->mu_s, O)); A_(dev_alloc(&p->sd_s, O)); A_(dev_alloc(&p->mu_a, A)); A_(dev_alloc(&p->sd_a, A));
  A_(dev_alloc(&p->cost_w, O)); A_(dev_alloc(&p->goal, O));
  A_(dev_alloc(&p->d_s0, (size_t)E * O)); A_(dev_alloc(&p->d_costs, (size_t)p->R));
  A_(dev_alloc(&p->d_mu_hist, EHA * (cfg->max_iterations + 1)));
  A_(dev_alloc(&p->d_sd_hist, EHA * (cfg->max_iterations + 1)));
  A_(dev_alloc(&p->d_mu_last, EHA));
  A_(dev_alloc(&p->d_elite, (size_t)E * cfg->max_elites));
  {
    // per slot: one partial sum per 2048-elite chunk, or -- single-environment planners, which may be
    // population-sharded -- per rank / refit segment (up to 64)
    const size_t slots = (size_t)E * H * ((A + 3) / 4);
    size_t chunks = (size_t)(cfg->max_elites + kRefitChunk - 1) / kRefitChunk;
    if (E == 1 && chunks < 64) chunks = 64;
    p->refit_parts_cap = (int)chunks;
    A_(dev_alloc(&p->d_own_count, 1));
    A_(dev_alloc(&p->d_refit_part, slots * chunks * 8));
    A_(dev_alloc(&p->d_refit_arrive, slots));
    if (ok) A_(cudaMemset(p->d_refit_arrive, 0, sizeof(unsigned int) * slots));
  }
  A_(dev_alloc(&p->d_best_ever, E)); A_(dev_alloc(&p->d_info, E));
  A_(dev_alloc(&p->d_out_states, (size_t)E * H * O)); A_(dev_alloc(&p->d_out_actions, EHA));
  A_(cudaMallocHost((void**)&p->h_s0, sizeof(float) * E * O));
  A_(cudaMallocHost((void**)&p->h_out_states, sizeof(float) * E * H * O));
  A_(cudaMallocHost((void**)&p->h_out_actions, sizeof(float) * EHA));
  A_(cudaMallocHost((void**)&p->h_mu, sizeof(float) * EHA));
  A_(cudaMallocHost((void**)&p->h_sd, sizeof(float) * EHA));
  A_(cudaMallocHost((void**)&p->h_info, sizeof(MbrlPlanInfo) * E));
  A_(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  if (ok && getenv("MBRL_PLAN_TIMELINE") && (p->R + kTcRows - 1) / kTcRows <= kPlanStampCtas) {
    const size_t n = (size_t)kPlanStampStride * cfg->max_iterations;
    if (cudaMalloc((void**)&p->d_plan_stamps, sizeof(long long) * n) == cudaSuccess) cudaMemset(p->d_plan_stamps, 0, sizeof(long long) * n);
    else { p->d_plan_stamps = nullptr; cudaGetLastError(); }
  }
  if (ok) {
    // identity normalisers until mbrl_set_norm is called
    std::vector<float> zeros(kMaxObs, 0.f), ones(kMaxObs, 1.f);
    A_(cudaMemcpy(p->mu_s, zeros.data(), sizeof(float) * O, cudaMemcpyHostToDevice));
    A_(cudaMemcpy(p->sd_s, ones.data(), sizeof(float) * O, cudaMemcpyHostToDevice));
    A_(cudaMemcpy(p->mu_a, zeros.data(), sizeof(float) * A, cudaMemcpyHostToDevice));
    A_(cudaMemcpy(p->sd_a, ones.data(), sizeof(float) * A, cudaMemcpyHostToDevice));
  }
  if (ok && cfg->engine != MBRL_ENGINE_SIMT_FP32) {
    std::string why;
    if (!tc_init(&p->tc, O, A, U, cfg->engine == MBRL_ENGINE_TC_FP16, p->max_smem, &why)) {
      mbrl_destroy(p);
      return fail(MBRL_E_UNSUPPORTED, "tensor-core engine: " + why);
    }
  }
  if (!ok) {
    std::string msg = g_err;
    mbrl_destroy(p);
    return fail(MBRL_E_CUDA, "mbrl_create: " + msg);
  }
  *out = p;
  return MBRL_OK;
}

// nn.Linear [out,in] row-major  ->  K-major [in][out]
static void transpose_to(std::vector<float>& dst, const float* W, int out_f, int in_f) {
  dst.resize((size_t)out_f * in_f);
  for (int o = 0; o < out_f; ++o)
    for (int i = 0; i < in_f; ++i) dst[(size_t)i * out_f + o] = W[(size_t)o * in_f + i];
}

extern "C" int mbrl_set_weights(MbrlPlanner* p, const float* W1, const float* b1, const float* W2,
                                const float* b2, const float* W3, const float* b3) {
  if (!p) return fail(MBRL_E_INVALID, "null planner");
  MBRL_REQUIRE(W1 && b1 && W2 && b2 && W3 && b3, "null weight pointer");
  MBRL_CUDA(cudaSetDevice(p->cfg.device));
  MBRL_CUDA(cudaDeviceSynchronize());  // a plan may be in flight on ANY stream (mbrl_plan_device runs on the caller's)
  std::vector<float> t;
  transpose_to(t, W1, p->U, p->D);
  MBRL_CUDA(cudaMemcpy(p->W1t, t.data(), sizeof(float) * t.size(), cudaMemcpyHostToDevice));
  transpose_to(t, W2, p->U, p->U);
  MBRL_CUDA(cudaMemcpy(p->W2t, t.data(), sizeof(float) * t.size(), cudaMemcpyHostToDevice));
  transpose_to(t, W3, p->O, p->U);
  MBRL_CUDA(cudaMemcpy(p->W3t, t.data(), sizeof(float) * t.size(), cudaMemcpyHostToDevice));
  MBRL_CUDA(cudaMemcpy(p->b1, b1, sizeof(float) * p->U, cudaMemcpyHostToDevice));
  MBRL_CUDA(cudaMemcpy(p->b2, b2, sizeof(float) * p->U, cudaMemcpyHostToDevice));
  MBRL_CUDA(cudaMemcpy(p->b3, b3, sizeof(float) * p->O, cudaMemcpyHostToDevice));
  if (p->cfg.engine != MBRL_ENGINE_SIMT_FP32) {
    std::string why;
    if (!tc_set_weights(&p->tc, W1, b1, W2, b2, W3, b3, &why)) return fail(MBRL_E_CUDA, why);
    p->hW1.assign(W1, W1 + (size_t)p->U * p->D); p->hb1.assign(b1, b1 + p->U);
    p->hW2.assign(W2, W2 + (size_t)p->U * p->U); p->hb2.assign(b2, b2 + p->U);
    p->hW3.assign(W3, W3 + (size_t)p->O * p->U); p->hb3.assign(b3, b3 + p->O);
  }
  p->have_weights = true;
  return MBRL_OK;
}

extern "C" int mbrl_set_norm(MbrlPlanner* p, const float* mu_s, const float* sd_s, const float* mu_a,
                             const float* sd_a) {
  if (!p) return fail(MBRL_E_INVALID, "null planner");
  MBRL_CUDA(cudaSetDevice(p->cfg.device));
  MBRL_CUDA(cudaDeviceSynchronize());  // a plan may be in flight on ANY stream (mbrl_plan_device runs on the caller's)
  std::vector<float> zeros(kMaxObs, 0.f), ones(kMaxObs, 1.f);
  MBRL_CUDA(cudaMemcpy(p->mu_s, mu_s ? mu_s : zeros.data(), sizeof(float) * p->O, cudaMemcpyHostToDevice));
  MBRL_CUDA(cudaMemcpy(p->sd_s, sd_s ? sd_s : ones.data(), sizeof(float) * p->O, cudaMemcpyHostToDevice));
  MBRL_CUDA(cudaMemcpy(p->mu_a, mu_a ? mu_a : zeros.data(), sizeof(float) * p->A, cudaMemcpyHostToDevice));
  MBRL_CUDA(cudaMemcpy(p->sd_a, sd_a ? sd_a : ones.data(), sizeof(float) * p->A, cudaMemcpyHostToDevice));
  return MBRL_OK;
}

extern "C" int mbrl_set_cost(MbrlPlanner* p, int32_t kind, const float* w, const float* goal, double alpha,
                             double beta) {
  if (!p) return fail(MBRL_E_INVALID, "null planner");
  MBRL_REQUIRE(kind == MBRL_COST_SMOOTHABS_COSH || kind == MBRL_COST_REWARD_HEAD || is_task_cost(kind), "unknown cost kind");
  MBRL_CUDA(cudaSetDevice(p->cfg.device));
  MBRL_CUDA(cudaDeviceSynchronize());  // a plan may be in flight on ANY stream (mbrl_plan_device runs on the caller's)
  if (is_task_cost(kind)) {
    int pick[4];
    task_pick_indices(kind, pick);
    const int need = std::max(std::max(pick[0], pick[1]), std::max(pick[2], pick[3])) + 1;
    MBRL_REQUIRE(p->O >= need && p->A >= 1, "observation too short for this dm_control task cost "
                 "(cartpole 5, cheetah 9, walker 17, humanoid 39 leading entries are read)");
    if (p->cfg.engine != MBRL_ENGINE_SIMT_FP32 && !tc_supports_task_cost(&p->tc, kind))
      return fail(MBRL_E_UNSUPPORTED, "the dm_control task-cost epilogue is not implemented for this tensor-core kernel variant");
    if (beta == 0.0) beta = 1.0;  // unused by these costs; keeps 1/beta finite
  } else if (kind == MBRL_COST_REWARD_HEAD) {
    MBRL_REQUIRE(p->have_reward_head, "mbrl_set_reward_head must be called before selecting the reward-head cost");
    if (p->cfg.engine != MBRL_ENGINE_SIMT_FP32 && !p->tc.head) {
      // RewardAgent's second trunk pass + linear4 head live in the weight-streaming kernel: switch this
      // handle to it (hidden <= 512) and re-pack the operand image
      std::string why;
      TcModel fresh{};
      if (!tc_init(&fresh, p->O, p->A, p->U, p->cfg.engine == MBRL_ENGINE_TC_FP16, p->max_smem, &why, true))
        return fail(MBRL_E_UNSUPPORTED, "reward-head cost on the tensor-core engine: " + why);
      fresh.d_dbg = p->tc.d_dbg; p->tc.d_dbg = nullptr;
      tc_free(&p->tc);
      p->tc = fresh;
      if (p->have_weights && !tc_set_weights(&p->tc, p->hW1.data(), p->hb1.data(), p->hW2.data(), p->hb2.data(), p->hW3.data(),
                                             p->hb3.data(), &why))
        return fail(MBRL_E_CUDA, why);
    }
    if (beta == 0.0) beta = 1.0;  // unused by this cost; keeps 1/beta finite
  } else {
    MBRL_REQUIRE(w && goal, "null cost pointer");
    MBRL_REQUIRE(beta != 0.0, "beta must be non-zero");
    MBRL_CUDA(cudaMemcpy(p->cost_w, w, sizeof(float) * p->O, cudaMemcpyHostToDevice));
    MBRL_CUDA(cudaMemcpy(p->goal, goal, sizeof(float) * p->O, cudaMemcpyHostToDevice));
  }
  p->cost_kind = kind;
  p->alpha = (float)alpha; p->alpha2 = (float)(alpha * alpha);
  p->beta = (float)beta; p->beta2 = (float)(beta * beta);
  p->have_cost = true;
  return MBRL_OK;
}

extern "C" int mbrl_set_reward_head(MbrlPlanner* p, const float* h_W4, float b4, float reward_mean, float reward_std) {
  if (!p) return fail(MBRL_E_INVALID, "null planner");
  MBRL_REQUIRE(h_W4, "null reward-head weights");
  MBRL_CUDA(cudaSetDevice(p->cfg.device));
  MBRL_CUDA(cudaDeviceSynchronize());  // a plan may be in flight on ANY stream (mbrl_plan_device runs on the caller's)
  if (!p->W4) MBRL_CUDA(dev_alloc(&p->W4, (size_t)p->U));
  MBRL_CUDA(cudaMemcpy(p->W4, h_W4, sizeof(float) * p->U, cudaMemcpyHostToDevice));
  p->b4 = b4; p->mu_r = reward_mean; p->sd_r = reward_std;
  p->have_reward_head = true;
  return MBRL_OK;
}

extern "C" int mbrl_set_action_bounds(MbrlPlanner* p, float lo, float hi) {
  if (!p) return fail(MBRL_E_INVALID, "null planner");
  MBRL_REQUIRE(lo <= hi, "lo > hi");
  p->lo = lo; p->hi = hi;
  return MBRL_OK;
}

// --------------------------------------------------------------------------------------
// kernel launchers
// --------------------------------------------------------------------------------------
// Launch with programmatic stream serialization (PDL): the kernel may begin while its stream
// predecessor is still running; every kernel of this library calls pdl_wait() before it touches
// data a predecessor produces (see common.cuh).
static bool g_use_pdl = getenv("MBRL_NO_PDL") == nullptr;
template <class... KArgs, class... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <int TM, int CPT, int CG>
static cudaError_t launch_simt_t(const MbrlPlanner* p, const ActionSource& src, const float* d_s0,
                                 float* d_costs, float* d_states, float* d_actions, cudaStream_t st) {
  const size_t smem = simt_smem_bytes<TM>(p->O, p->A, p->U);
  auto kern = rollout_simt_kernel<TM, CPT, CG>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  Shape sh{p->H, p->N, p->E};
  const unsigned grid = (unsigned)((p->R + TM - 1) / TM);
  kern<<<grid, TM * 4 * CG, smem, st>>>(model_view(p), src, sh, d_s0, d_costs, d_states, d_actions);
  return cudaGetLastError();
}

template <int TM>
static cudaError_t launch_simt_cpt(const MbrlPlanner* p, const ActionSource& src, const float* d_s0,
                                   float* d_costs, float* d_states, float* d_actions, cudaStream_t st) {
  const int U = p->U;
  // (columns per lane CPT, column groups CG): a pass covers 32*CPT*CG output columns.  Up to 256 hidden
  // units one warp per 8-row strip with a wide register tile is fastest (cfg 3: 2.76 ms vs 4.31 ms with
  // two column groups); wider layers shrink the row tile (shared memory) to 4 warps per CTA, and
  // splitting the columns over CG warps per strip restores the occupancy (cfg 5: 524 ms vs 928 ms).
  constexpr int kMaxCG = 1024 / (TM * 4);
  static const bool narrow = std::getenv("MBRL_SIMT_CG1") != nullptr;  // A/B switch: always one warp per strip
  if (U <= 64) return launch_simt_t<TM, 2, 1>(p, src, d_s0, d_costs, d_states, d_actions, st);
  if (U <= 128) return launch_simt_t<TM, 4, 1>(p, src, d_s0, d_costs, d_states, d_actions, st);
  if (U <= 224) return launch_simt_t<TM, 7, 1>(p, src, d_s0, d_costs, d_states, d_actions, st);
  if (U <= 256 || narrow) return launch_simt_t<TM, 8, 1>(p, src, d_s0, d_costs, d_states, d_actions, st);
  return launch_simt_t<TM, 4, (kMaxCG >= 4 ? 4 : kMaxCG)>(p, src, d_s0, d_costs, d_states, d_actions, st);
}

static int launch_rollout(MbrlPlanner* p, const ActionSource& src, const float* d_s0, float* d_costs,
                          float* d_states, float* d_actions, cudaStream_t st) {
  if (!p->have_weights || !p->have_cost)
    return fail(MBRL_E_STATE, "mbrl_set_weights and mbrl_set_cost must be called before planning");
  if ((src.mode == MBRL_SAMPLE_INJECT_ACTIONS || src.mode == MBRL_SAMPLE_INJECT_NOISE) && !src.buf)
    return fail(MBRL_E_INVALID, "injected sample mode without an injected buffer");
  if ((src.mode == MBRL_SAMPLE_INJECT_NOISE || src.mode == MBRL_SAMPLE_GAUSSIAN) && (!src.mu || !src.sd))
    return fail(MBRL_E_INVALID, "Gaussian sample mode without mu/sd");
  if (src.mode < MBRL_SAMPLE_INJECT_ACTIONS || src.mode > MBRL_SAMPLE_UNIFORM)
    return fail(MBRL_E_INVALID, "unknown sample mode");
  cudaError_t e;
  if (p->cfg.engine == MBRL_ENGINE_SIMT_FP32) {
    if (simt_smem_bytes<64>(p->O, p->A, p->U) <= p->max_smem)
      e = launch_simt_cpt<64>(p, src, d_s0, d_costs, d_states, d_actions, st);
    else if (simt_smem_bytes<32>(p->O, p->A, p->U) <= p->max_smem)
      e = launch_simt_cpt<32>(p, src, d_s0, d_costs, d_states, d_actions, st);
    else if (simt_smem_bytes<16>(p->O, p->A, p->U) <= p->max_smem)
      e = launch_simt_cpt<16>(p, src, d_s0, d_costs, d_states, d_actions, st);
    else
      return fail(MBRL_E_UNSUPPORTED, "model too large for the fp32 engine's shared-memory tiling");
  } else {
    Shape sh{p->H, p->N, p->E};
    e = tc_launch_rollout(&p->tc, model_view(p), src, sh, d_s0, d_costs, d_states, d_actions, p->num_sms, st);
  }
  if (e != cudaSuccess) return fail(MBRL_E_CUDA, std::string("rollout launch: ") + cudaGetErrorString(e));
  return MBRL_OK;
}

template <int MODE>
static int launch_topk_mode(const float* d_costs, int segments, int n, int k, int* d_idx, float* d_cost,
                            MbrlPlanInfo* d_best, BestEver* d_best_ever, int iteration, const SelShard& sh, cudaStream_t st) {
  MBRL_REQUIRE(segments >= 1 && n >= 1, "topk: empty input");
  MBRL_REQUIRE(k >= 1 && k <= n, "topk: k out of range [1,n]");
  if (n <= kSelectStageMax) {
    const size_t smem = sizeof(uint32_t) * (size_t)select_padded(n);
    MBRL_CUDA(cudaFuncSetAttribute(topk_select_kernel<true, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MBRL_CUDA(launch_pdl(topk_select_kernel<true, MODE>, dim3(segments), dim3(kSelectThreads), smem, st, d_costs, n, k, d_idx,
                         d_cost, d_best, d_best_ever, iteration, sh));
  } else {
    const size_t smem = sizeof(uint32_t) * 16 * kSelectThreads;  // compaction buffer of one 16384-key pass
    MBRL_CUDA(cudaFuncSetAttribute(topk_select_kernel<false, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MBRL_CUDA(launch_pdl(topk_select_kernel<false, MODE>, dim3(segments), dim3(kSelectThreads), smem, st, d_costs, n, k, d_idx,
                         d_cost, d_best, d_best_ever, iteration, sh));
  }
  MBRL_CUDA(cudaGetLastError());
  return MBRL_OK;
}

static int launch_topk(const float* d_costs, int segments, int n, int k, int* d_idx, float* d_cost,
                       MbrlPlanInfo* d_best, BestEver* d_best_ever, int iteration, cudaStream_t st, long long* stamps = nullptr) {
  SelShard none{};
  none.stamps = stamps;
  return launch_topk_mode<kSelPlain>(d_costs, segments, n, k, d_idx, d_cost, d_best, d_best_ever, iteration, none, st);
}

static int launch_refit(const MbrlPlanner* p, const ActionSource& src, const int* d_elite, int k,
                        float* d_mu_new, float* d_sd_new, cudaStream_t st) {
  MBRL_REQUIRE(k >= 1 && k <= p->cfg.max_elites, "refit: k out of range [1, max_elites]");
  const int G = (p->A + 3) / 4;
  Shape sh{p->H, p->N, p->E};
  dim3 grid(p->H * G, p->E, (k + kRefitChunk - 1) / kRefitChunk);
  const int threads = refit_threads(k);
  if (threads <= kRefitThreads / 2)  // register-capped build: three CTAs per SM
    MBRL_CUDA(launch_pdl(refit_kernel<kRefitThreads / 2, 3>, grid, dim3(threads), 0, st, src, sh, p->A, d_elite, k, d_mu_new,
                         d_sd_new, p->d_refit_part, p->d_refit_arrive, p->cur_stamps));
  else
    MBRL_CUDA(launch_pdl(refit_kernel<kRefitThreads, 1>, grid, dim3(threads), 0, st, src, sh, p->A, d_elite, k, d_mu_new,
                         d_sd_new, p->d_refit_part, p->d_refit_arrive, p->cur_stamps));
  MBRL_CUDA(cudaGetLastError());
  return MBRL_OK;
}

// Segment-canonical refit from the global elite list (see refit_seg_kernel): nseg segments of seg_size candidates.
static int launch_refit_seg(const MbrlPlanner* p, const ActionSource& src, const int* d_elite, int k, int nseg, int seg_size,
                            float* d_mu_new, float* d_sd_new, cudaStream_t st) {
  MBRL_REQUIRE(k >= 1 && k <= p->cfg.max_elites, "refit: k out of range [1, max_elites]");
  MBRL_REQUIRE(nseg >= 1 && nseg <= p->refit_parts_cap, "refit: too many segments");
  const int G = (p->A + 3) / 4;
  Shape sh{p->H, p->N, p->E};
  MBRL_CUDA(launch_pdl(refit_seg_kernel, dim3(p->H * G, p->E, nseg), dim3(kRefitThreads), 0, st, src, sh, p->A, d_elite, k,
                       seg_size, d_mu_new, d_sd_new, p->d_refit_part, p->d_refit_arrive));
  MBRL_CUDA(cudaGetLastError());
  return MBRL_OK;
}

static int launch_replay(MbrlPlanner* p, int mode, uint64_t seed, uint32_t cand_offset, uint32_t env_offset,
                         const float* d_s0, const float* d_injected, const float* d_mu_hist,
                         const float* d_sd_hist, int iterations, int return_mean, int actions_only,
                         const BestEver* d_best, float* d_out_states, float* d_out_actions, MbrlPlanInfo* d_info,
                         cudaStream_t st) {
  if (!p->have_weights) return fail(MBRL_E_STATE, "mbrl_set_weights must be called before planning");
  ActionSource src = action_source(p, mode, seed, 0, cand_offset, env_offset, d_injected, d_mu_hist, d_sd_hist);
  Shape sh{p->H, p->N, p->E};
  const size_t smem_w = replay_smem_bytes(p->O, p->A, p->U, p->H, true);
  const RegGeom rg = replay_reg_geometry(p->O, p->A, p->U, p->H);
  static const bool no_reg = std::getenv("MBRL_REPLAY_NO_REG") != nullptr;  // A/B switch for profiling
  if (rg.ok && sizeof(float) * (size_t)rg.total <= p->max_smem && !actions_only && !no_reg) {
    const size_t smem = sizeof(float) * (size_t)rg.total;
    auto launch = [&](auto kernel) -> cudaError_t {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      return launch_pdl(kernel, dim3(p->E), dim3(kRegThreads), smem, st, model_view(p), src, sh, rg, d_s0, d_mu_hist,
                        d_sd_hist, d_best, iterations, return_mean, d_out_states, d_out_actions, d_info);
    };
    if (rg.kpt2 == 16) MBRL_CUDA(launch(replay_reg_kernel<8, 16>));
    else MBRL_CUDA(launch(replay_reg_kernel<5, 40>));
  } else if (smem_w <= p->max_smem && !actions_only) {
    MBRL_CUDA(cudaFuncSetAttribute(replay_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
    MBRL_CUDA(launch_pdl(replay_kernel<true>, dim3(p->E), dim3(kReplayThreads), smem_w, st, model_view(p), src, sh, d_s0,
                         d_mu_hist, d_sd_hist, d_best, iterations, return_mean, actions_only, d_out_states, d_out_actions, d_info));
  } else {
    const size_t smem = replay_smem_bytes(p->O, p->A, p->U, p->H, false);
    MBRL_REQUIRE(smem <= p->max_smem, "horizon too long for the replay kernel's shared memory");
    MBRL_CUDA(cudaFuncSetAttribute(replay_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    replay_kernel<false><<<p->E, kReplayThreads, smem, st>>>(model_view(p), src, sh, d_s0, d_mu_hist, d_sd_hist, d_best,
                                                            iterations, return_mean, actions_only, d_out_states, d_out_actions, d_info);
  }
  MBRL_CUDA(cudaGetLastError());
  return MBRL_OK;
}

// --------------------------------------------------------------------------------------
// building-block entry points
// --------------------------------------------------------------------------------------
extern "C" int mbrl_rollout(MbrlPlanner* p, int32_t mode, uint64_t seed, uint32_t iteration,
                            uint32_t cand_offset, uint32_t env_offset, const float* d_s0,
                            const float* d_injected, const float* d_mu, const float* d_sd, float* d_costs,
                            float* d_states_out, float* d_actions_out, void* stream) {
  if (!p) return fail(MBRL_E_INVALID, "null planner");
  MBRL_REQUIRE(d_s0 && d_costs, "null s0/costs");
  MBRL_CUDA(cudaSetDevice(p->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;  // NULL = the CUDA default stream, as everywhere in CUDA
  ActionSource src = action_source(p, mode, seed, iteration, cand_offset, env_offset, d_injected, d_mu, d_sd);
  return launch_rollout(p, src, d_s0, d_costs, d_states_out, d_actions_out, st);
}

extern "C" int mbrl_sample(MbrlPlanner* p, int32_t mode, uint64_t seed, uint32_t iteration,
                           uint32_t cand_offset, uint32_t env_offset, const float* d_mu, const float* d_sd,
                           float* d_out, void* stream) {
  if (!p) return fail(MBRL_E_INVALID, "null planner");
  MBRL_REQUIRE(mode == MBRL_SAMPLE_GAUSSIAN || mode == MBRL_SAMPLE_UNIFORM, "sample: Philox modes only");
  MBRL_REQUIRE(d_out, "null output");
  MBRL_REQUIRE(mode == MBRL_SAMPLE_UNIFORM || (d_mu && d_sd), "Gaussian mode needs mu/sd");
  MBRL_CUDA(cudaSetDevice(p->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;  // NULL = the CUDA default stream, as everywhere in CUDA
  ActionSource src = action_source(p, mode, seed, iteration, cand_offset, env_offset, nullptr, d_mu, d_sd);
  Shape sh{p->H, p->N, p->E};
  const long long total = p->R * p->H;
  sample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, sh, p->A, d_out);
  MBRL_CUDA(cudaGetLastError());
  return MBRL_OK;
}

extern "C" int mbrl_philox_raw(const uint32_t* d_ctr, const uint32_t* d_key, uint32_t* d_out, int64_t n,
                               void* stream) {
  MBRL_REQUIRE(d_ctr && d_key && d_out && n >= 0, "philox_raw: bad argument");
  if (n == 0) return MBRL_OK;
  philox_raw_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_ctr, d_key, d_out, n);
  MBRL_CUDA(cudaGetLastError());
  return MBRL_OK;
}

extern "C" int mbrl_topk(const float* d_costs, int32_t segments, int32_t n, int32_t k, int32_t* d_elite_idx,
                         float* d_elite_cost, MbrlPlanInfo* d_best, void* stream) {
  MBRL_REQUIRE(d_costs && d_elite_idx, "topk: null pointer");
  return launch_topk(d_costs, segments, n, k, d_elite_idx, d_elite_cost, d_best, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int mbrl_refit(MbrlPlanner* p, int32_t mode, uint64_t seed, uint32_t iteration,
                          uint32_t cand_offset, uint32_t env_offset, const float* d_injected,
                          const float* d_mu, const float* d_sd, const int32_t* d_elite_idx, int32_t k,
                          float* d_mu_new, float* d_sd_new, void* stream) {
  if (!p) return fail(MBRL_E_INVALID, "null planner");
  MBRL_REQUIRE(d_elite_idx && d_mu_new && d_sd_new, "refit: null pointer");
  MBRL_REQUIRE(mode >= MBRL_SAMPLE_INJECT_ACTIONS && mode <= MBRL_SAMPLE_UNIFORM, "unknown sample mode");
  MBRL_REQUIRE((mode != MBRL_SAMPLE_INJECT_ACTIONS && mode != MBRL_SAMPLE_INJECT_NOISE) || d_injected,
               "injected mode without buffer");
  MBRL_REQUIRE((mode != MBRL_SAMPLE_INJECT_NOISE && mode != MBRL_SAMPLE_GAUSSIAN) || (d_mu && d_sd),
               "Gaussian mode needs mu/sd");
  MBRL_CUDA(cudaSetDevice(p->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;  // NULL = the CUDA default stream, as everywhere in CUDA
  ActionSource src = action_source(p, mode, seed, iteration, cand_offset, env_offset, d_injected, d_mu, d_sd);
  return launch_refit(p, src, d_elite_idx, k, d_mu_new, d_sd_new, st);
}

static_assert(sizeof(BestEver) == sizeof(MbrlPlanInfo), "BestEver and MbrlPlanInfo share one layout");

extern "C" int mbrl_nccl_unique_id(uint8_t* h_id128) {
  MBRL_REQUIRE(h_id128, "null id buffer");
  int rc = load_nccl();
  if (rc) return rc;
  NcclUniqueId id;
  MBRL_NCCL(g_nccl.GetUniqueId(&id));
  std::memcpy(h_id128, id.internal, 128);
  return MBRL_OK;
}

extern "C" int mbrl_comm_destroy(MbrlPlanner* p) {
  if (!p) return fail(MBRL_E_INVALID, "null planner");
  if (p->comm) {
    MBRL_CUDA(cudaSetDevice(p->cfg.device));
    MBRL_CUDA(cudaDeviceSynchronize());
    g_nccl.CommDestroy(p->comm);
    p->comm = nullptr;
  }
  p->rank = 0; p->world = 1;
  return MBRL_OK;
}

static unsigned long long p2p_timeout_ns() {
  static const unsigned long long ns = [] {
    const char* e = getenv("MBRL_P2P_TIMEOUT_S");
    const double sec = e ? std::atof(e) : 120.0;
    return (unsigned long long)((sec > 0.0 ? sec : 120.0) * 1e9);
  }();
  return ns;
}

static int alloc_shard_scratch(MbrlPlanner* p, int world) {
  if (p->scratch_world == world) return MBRL_OK;
  // (re)allocate for this world size; commit only when every allocation succeeded
  void* old[] = {p->d_ecost, p->d_send, p->d_recv, p->d_gcost, p->d_gidx, p->d_pos, p->d_best_now, p->d_trunc};
  for (void* q : old) if (q) cudaFree(q);
  p->d_ecost = nullptr; p->d_send = nullptr; p->d_recv = nullptr; p->d_gcost = nullptr; p->d_gidx = nullptr;
  p->d_pos = nullptr; p->d_best_now = nullptr; p->d_trunc = nullptr; p->scratch_world = 0;
  const size_t kl = (size_t)std::min(p->cfg.max_elites, p->N);
  bool ok = dev_alloc(&p->d_ecost, kl) == cudaSuccess && dev_alloc(&p->d_send, 2 * kl) == cudaSuccess &&
            dev_alloc(&p->d_recv, 2 * kl * world) == cudaSuccess && dev_alloc(&p->d_gcost, kl * world) == cudaSuccess &&
            dev_alloc(&p->d_gidx, kl * world) == cudaSuccess && dev_alloc(&p->d_pos, (size_t)p->cfg.max_elites) == cudaSuccess &&
            dev_alloc(&p->d_best_now, 1) == cudaSuccess && dev_alloc(&p->d_trunc, 1) == cudaSuccess &&
            cudaMemset(p->d_trunc, 0, sizeof(int)) == cudaSuccess;
  if (!ok) {
    void* part[] = {p->d_ecost, p->d_send, p->d_recv, p->d_gcost, p->d_gidx, p->d_pos, p->d_best_now, p->d_trunc};
    for (void* q : part) if (q) cudaFree(q);
    p->d_ecost = nullptr; p->d_send = nullptr; p->d_recv = nullptr; p->d_gcost = nullptr; p->d_gidx = nullptr;
    p->d_pos = nullptr; p->d_best_now = nullptr; p->d_trunc = nullptr;
    cudaGetLastError();
    return fail(MBRL_E_CUDA, "out of device memory allocating the population-sharding scratch buffers");
  }
  p->scratch_world = world;
  return MBRL_OK;
}

extern "C" int mbrl_p2p_export(MbrlPlanner* p, int32_t world, uint8_t* h_handle64) {
  if (!p) return fail(MBRL_E_INVALID, "null planner");
  MBRL_REQUIRE(h_handle64 && world >= 1 && world <= 64, "p2p_export: bad argument");
  MBRL_REQUIRE(p->E == 1, "population sharding needs num_envs == 1");
  MBRL_REQUIRE(!p->p2p_attached, "detach the peer buffers (mbrl_p2p_detach) before exporting again");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  MBRL_CUDA(cudaSetDevice(p->cfg.device));
  const int slot = std::min(p->cfg.max_elites, p->N);
  const size_t words = p2p_total_words(world, slot, p->H * ((p->A + 3) / 4));
  if (p->d_p2p_local && p->p2p_world != world) {  // another world size: a new buffer
    MBRL_CUDA(cudaDeviceSynchronize());
    cudaFree(p->d_p2p_local); p->d_p2p_local = nullptr;
  }
  if (!p->d_p2p_local) {
    uint32_t* buf = nullptr;
    int* err = p->d_p2p_error;
    bool ok = dev_alloc(&buf, words) == cudaSuccess && (err || dev_alloc(&err, 1) == cudaSuccess);
    if (!ok) {  // nothing half-initialised is left behind: a retry starts from scratch
      if (buf) cudaFree(buf);
      if (err && err != p->d_p2p_error) cudaFree(err);
      cudaGetLastError();
      return fail(MBRL_E_CUDA, "out of device memory allocating the peer-memory gather buffer");
    }
    p->d_p2p_local = buf; p->d_p2p_error = err;
  }
  // (re-)export: packet tags and the sequence number restart from zero on every rank
  p->p2p_slot = slot;
  p->p2p_world = world;
  p->p2p_seq = 0;
  MBRL_CUDA(cudaMemset(p->d_p2p_local, 0, sizeof(uint32_t) * words));
  MBRL_CUDA(cudaMemset(p->d_p2p_error, 0, sizeof(int)));
  cudaIpcMemHandle_t hdl;
  MBRL_CUDA(cudaIpcGetMemHandle(&hdl, p->d_p2p_local));
  std::memcpy(h_handle64, &hdl, 64);
  return MBRL_OK;
}

extern "C" int mbrl_p2p_attach(MbrlPlanner* p, const uint8_t* h_handles, int32_t rank, int32_t world) {
  if (!p) return fail(MBRL_E_INVALID, "null planner");
  MBRL_REQUIRE(h_handles && p->d_p2p_local && world == p->p2p_world && rank >= 0 && rank < world,
               "p2p_attach: export first, with the same world size");
  MBRL_REQUIRE(!p->p2p_attached, "p2p already attached");
  MBRL_CUDA(cudaSetDevice(p->cfg.device));
  if (const char* f = getenv("MBRL_TEST_P2P_FAIL_RANK"))  // tests: make one rank's attach fail (fallback path)
    if (std::atoi(f) == rank) return fail(MBRL_E_CUDA, "p2p_attach: forced failure (MBRL_TEST_P2P_FAIL_RANK)");
  for (int r = 0; r < world; ++r) p->p2p_peers.base[r] = nullptr;
  p->rank = rank; p->world = world;
  for (int r = 0; r < world; ++r) {
    if (r == rank) { p->p2p_peers.base[r] = p->d_p2p_local; continue; }
    cudaIpcMemHandle_t hdl;
    std::memcpy(&hdl, h_handles + 64 * r, 64);
    void* ptr = nullptr;
    MBRL_CUDA(cudaIpcOpenMemHandle(&ptr, hdl, cudaIpcMemLazyEnablePeerAccess));  // on failure: mbrl_p2p_detach cleans up
    p->p2p_peers.base[r] = (uint32_t*)ptr;
  }
  p->p2p_attached = true;
  if (getenv("MBRL_SHARD_TIMELINE") && !p->d_shard_stamps) {
    const size_t n = (size_t)16 * p->cfg.max_iterations;
    if (cudaMalloc((void**)&p->d_shard_stamps, sizeof(long long) * n) == cudaSuccess) cudaMemset(p->d_shard_stamps, 0, sizeof(long long) * n);
    else { p->d_shard_stamps = nullptr; cudaGetLastError(); }
  }
  return alloc_shard_scratch(p, world);
}

extern "C" int mbrl_set_refit_segments(MbrlPlanner* p, int32_t segments) {
  if (!p) return fail(MBRL_E_INVALID, "null planner");
  MBRL_REQUIRE(segments >= 1 && segments <= p->refit_parts_cap, "refit segments out of range [1, 64]");
  MBRL_REQUIRE(p->N % segments == 0, "refit segments must divide the population");
  MBRL_REQUIRE(p->comm == nullptr && !p->p2p_attached, "a population shard takes its segments from the world size");
  p->refit_segments = segments;
  return MBRL_OK;
}

extern "C" int mbrl_p2p_detach(MbrlPlanner* p) {
  if (!p) return fail(MBRL_E_INVALID, "null planner");
  MBRL_CUDA(cudaSetDevice(p->cfg.device));
  MBRL_CUDA(cudaDeviceSynchronize());
  for (int r = 0; r < p->p2p_world && r < 64; ++r) {
    if (r != p->rank && p->p2p_peers.base[r]) cudaIpcCloseMemHandle(p->p2p_peers.base[r]);
    p->p2p_peers.base[r] = nullptr;
  }
  p->p2p_attached = false;
  return MBRL_OK;
}

extern "C" int mbrl_comm_init(MbrlPlanner* p, const uint8_t* h_id128, int32_t rank, int32_t world) {
  if (!p) return fail(MBRL_E_INVALID, "null planner");
  MBRL_REQUIRE(h_id128, "null id");
  MBRL_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, "bad rank/world");
  MBRL_REQUIRE(p->E == 1, "population sharding needs num_envs == 1 (shard environments without a communicator)");
  MBRL_REQUIRE(!p->comm, "communicator already initialised");
  int rc = load_nccl();
  if (rc) return rc;
  MBRL_CUDA(cudaSetDevice(p->cfg.device));
  NcclUniqueId id;
  std::memcpy(id.internal, h_id128, 128);
  MBRL_NCCL(g_nccl.CommInitRank(&p->comm, world, id, rank));
  MBRL_REQUIRE(!p->p2p_attached || (p->rank == rank && p->world == world), "rank/world differ from the p2p attachment");
  p->rank = rank; p->world = world;
  return alloc_shard_scratch(p, world);
}

extern "C" int mbrl_emit(MbrlPlanner* p, int32_t mode, uint64_t seed, uint32_t cand_offset, uint32_t env_offset,
                         const float* d_s0, const float* d_injected, const float* d_mu_hist,
                         const float* d_sd_hist, int32_t iterations, int32_t return_mean,
                         const MbrlPlanInfo* d_best, float* d_out_states, float* d_out_actions, void* stream) {
  if (!p) return fail(MBRL_E_INVALID, "null planner");
  MBRL_REQUIRE(d_s0 && d_best && d_out_states && d_out_actions, "emit: null pointer");
  MBRL_REQUIRE(d_mu_hist && d_sd_hist, "emit: null mu/sd history");
  MBRL_REQUIRE(iterations >= 1, "emit: iterations must be >= 1");
  MBRL_CUDA(cudaSetDevice(p->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;  // NULL = the CUDA default stream, as everywhere in CUDA
  return launch_replay(p, mode, seed, cand_offset, env_offset, d_s0, d_injected, d_mu_hist, d_sd_hist, iterations,
                       return_mean, 0, reinterpret_cast<const BestEver*>(d_best), d_out_states, d_out_actions, nullptr, st);
}

extern "C" int mbrl_tc_debug(MbrlPlanner* p, int32_t enable, float* h_out) {
  if (!p) return fail(MBRL_E_INVALID, "null planner");
  MBRL_REQUIRE(p->cfg.engine != MBRL_ENGINE_SIMT_FP32, "tc_debug: not a tensor-core engine");
  MBRL_CUDA(cudaSetDevice(p->cfg.device));
  const size_t bytes = sizeof(float) * kTcDbgFloats + sizeof(long long) * kTcTimelineSteps * kTcTimelineEvents;
  if (h_out) {
    MBRL_REQUIRE(p->tc.d_dbg, "tc_debug: dump was never armed");
    MBRL_CUDA(cudaDeviceSynchronize());
    MBRL_CUDA(cudaMemcpy(h_out, p->tc.d_dbg, bytes, cudaMemcpyDeviceToHost));
  }
  if (enable && !p->tc.d_dbg) {
    MBRL_CUDA(cudaMalloc((void**)&p->tc.d_dbg, bytes));
    MBRL_CUDA(cudaMemset(p->tc.d_dbg, 0, bytes));
  } else if (!enable && p->tc.d_dbg) {
    MBRL_CUDA(cudaDeviceSynchronize());
    MBRL_CUDA(cudaFree(p->tc.d_dbg));
    p->tc.d_dbg = nullptr;
  }
  return MBRL_OK;
}

// --------------------------------------------------------------------------------------
// whole plans
// --------------------------------------------------------------------------------------
static int enqueue_plan(MbrlPlanner* p, const MbrlPlanArgs* a, const float* d_s0, const float* d_injected,
                        float* d_out_states, float* d_out_actions, MbrlPlanInfo* d_info, bool need_final_dist,
                        cudaStream_t st) {
  const int I = a->iterations;
  MBRL_REQUIRE(I >= 1 && I <= p->cfg.max_iterations, "iterations out of range [1, max_iterations]");
  need_final_dist = need_final_dist || (a->warm_start & MBRL_WARM_KEEP) != 0;  // the kept mean is the refit after the last iteration
  const int k_all = (I == 1 && !need_final_dist) ? 1 : a->elites;
  MBRL_REQUIRE(k_all >= 1 && k_all <= p->cfg.max_elites, "elites out of range [1, max_elites]");
  const size_t EHA = (size_t)p->E * p->H * p->A;
  const long long HRA = (long long)p->H * p->R * p->A;

  MBRL_REQUIRE((a->warm_start & ~(MBRL_WARM_USE | MBRL_WARM_KEEP)) == 0, "unknown warm_start bits");
  const bool warm_keep = (a->warm_start & MBRL_WARM_KEEP) != 0;
  if (a->h_mu0 && a->h_sd0) {
    MBRL_CUDA(cudaMemcpyAsync(p->d_mu_hist, a->h_mu0, sizeof(float) * EHA, cudaMemcpyHostToDevice, st));
    MBRL_CUDA(cudaMemcpyAsync(p->d_sd_hist, a->h_sd0, sizeof(float) * EHA, cudaMemcpyHostToDevice, st));
    init_plan_kernel<<<(unsigned)((p->E + 255) / 256), 256, 0, st>>>(nullptr, nullptr, 0, p->lo, p->hi, p->d_best_ever, p->E, p->stage_s0);
  } else if ((a->warm_start & MBRL_WARM_USE) && p->have_last) {
    // device-resident warm start: no host round trip of the mean between MPC steps
    MBRL_REQUIRE(!a->h_mu0 && !a->h_sd0, "mu0 and sd0 must be given together");
    const float std = a->warm_std > 0.f ? a->warm_std : 0.5f * (p->hi - p->lo);
    const long long n = (long long)EHA > p->E ? (long long)EHA : p->E;
    warm_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p->d_mu_hist, p->d_sd_hist, p->d_mu_last, p->E, p->H, p->A, std,
                                                                  p->lo, p->hi, p->d_best_ever, p->stage_s0);
  } else {
    MBRL_REQUIRE(!a->h_mu0 && !a->h_sd0, "mu0 and sd0 must be given together");
    const long long n = (long long)EHA > p->E ? (long long)EHA : p->E;
    init_plan_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p->d_mu_hist, p->d_sd_hist, (long long)EHA, p->lo, p->hi, p->d_best_ever, p->E, p->stage_s0);
  }
  MBRL_CUDA(cudaGetLastError());

  const bool sharded = p->comm != nullptr || p->p2p_attached;
  if (p->p2p_attached) MBRL_CUDA(cudaMemsetAsync(p->d_p2p_error, 0, sizeof(int), st));  // one late rank must not poison later plans
  if (sharded) {
    MBRL_REQUIRE(a->sample_mode == MBRL_SAMPLE_GAUSSIAN || a->sample_mode == MBRL_SAMPLE_UNIFORM,
                 "population sharding supports the Philox sample modes only");
    MBRL_REQUIRE((long long)k_all <= (long long)p->world * p->N, "elites exceed the global population");
  } else {
    MBRL_REQUIRE(k_all <= p->N, "elites exceed the population");
  }
  const uint32_t cand_offset = sharded ? (uint32_t)((long long)p->rank * p->N) : a->cand_offset;
  for (int it = 0; it < I; ++it) {
    // the last iteration of a plan that keeps no distribution only needs its best candidate: a k = 1 select
    const int k = (it + 1 == I && !need_final_dist) ? 1 : k_all;
    const float* inj = d_injected ? d_injected + (long long)it * HRA : nullptr;
    ActionSource src = action_source(p, a->sample_mode, a->seed, (uint32_t)it, cand_offset, a->env_offset, inj,
                                     p->d_mu_hist + it * EHA, p->d_sd_hist + it * EHA);
    p->cur_stamps = (p->d_plan_stamps && !sharded) ? p->d_plan_stamps + (size_t)kPlanStampStride * it : nullptr;
    p->tc.fg.stamps = p->cur_stamps ? p->cur_stamps + 8 : nullptr;
    int rc = launch_rollout(p, src, d_s0, p->d_costs, nullptr, nullptr, st);
    if (rc) return rc;
    if (sharded) {
      // local elites -> all-gather (cost, global index) -> same global top-k threshold on every rank ->
      // refit from GLOBAL indices (cand_offset 0) as per-rank partial sums added in rank order: over
      // peer memory every rank sums its own elites and the partials are exchanged, over NCCL every rank
      // holds the whole list and computes all partials itself (no second collective) -- same bits
      // Worst case a single shard holds all k global elites (k_full = min(k, N) per rank); the
      // shards are i.i.d., so a rank's share is Binomial(k, 1/world): send the expected share plus
      // 8 standard deviations (+64) and verify exactness on the device (the merge select); a
      // flagged plan -- probability ~1e-15 per iteration -- is redone with full-size gathers.
      const int k_full = std::min(k, p->N);
      const int share = (k + p->world - 1) / p->world;
      int kl = (p->full_gather || p->world == 1) ? k_full : std::min(k_full, share + 8 * (int)std::ceil(std::sqrt((double)share)) + 64);
      if (const char* force = getenv("MBRL_SHARD_KL")) {  // tests: "min" = the smallest legal gather (exactly the
        // expected share), which the exactness check must flag so that the redo path is exercised
        if (!p->full_gather && p->world > 1 && force[0] == 'm') kl = std::min(k_full, (k + p->world - 1) / p->world);
      }
      const int ng = kl * p->world;
      if (p->p2p_attached) {
        // peer-memory exchange: the local select stores its elites as sequence-tagged packets straight into
        // every rank's buffer; the merge select (resident and polling already) stages them as they land,
        // finds the global threshold and keeps this rank's own elites; the refit below is distributed
        SelShard sh{};
        sh.peers = p->p2p_peers; sh.local = p->d_p2p_local; sh.rank = p->rank; sh.world = p->world;
        sh.slot = p->p2p_slot; sh.pslots = p->H * ((p->A + 3) / 4); sh.seq = ++p->p2p_seq; sh.parity = (int)(sh.seq & 1);
        sh.idx_offset = (int)cand_offset; sh.k_full = k_full; sh.trunc = p->d_trunc; sh.error = p->d_p2p_error;
        sh.own_count = p->d_own_count; sh.timeout_ns = p2p_timeout_ns();
        sh.stamps = p->d_shard_stamps ? p->d_shard_stamps + 16 * it : nullptr;
        rc = launch_topk_mode<kSelScatter>(p->d_costs, 1, p->N, kl, nullptr, nullptr, nullptr, nullptr, it, sh, st);
        if (rc) return rc;
        rc = launch_topk_mode<kSelMerge>(nullptr, 1, ng, k, p->d_elite, nullptr, nullptr, p->d_best_ever, it, sh, st);
        if (rc) return rc;
        if (it + 1 < I || need_final_dist) {
          // distributed refit: this rank's elites only, partial sums exchanged through peer memory
          ActionSource gsrc = src;
          gsrc.cand_offset = 0;  // the merge emitted global candidate indices
          float* mu_new = p->d_mu_hist + (it + 1) * EHA;
          float* sd_new = p->d_sd_hist + (it + 1) * EHA;
          if (p->world == 1) {   // the only rank holds every elite: the plain refit, bit for bit
            rc = launch_refit(p, gsrc, p->d_elite, k, mu_new, sd_new, st);
            if (rc) return rc;
          } else {
            RefitP2p px{};
            px.peers = p->p2p_peers; px.local = p->d_p2p_local; px.rank = p->rank; px.world = p->world; px.slot = p->p2p_slot;
            px.pslots = sh.pslots; px.parity = sh.parity; px.seq = sh.seq; px.error = p->d_p2p_error; px.timeout_ns = sh.timeout_ns;
            px.stamps = sh.stamps;
            Shape shp{p->H, p->N, p->E};
            MBRL_CUDA(launch_pdl(refit_p2p_kernel, dim3(sh.pslots), dim3(kRefitThreads), 0, st, gsrc, shp, p->A,
                                 (const int*)p->d_elite, (const int*)p->d_own_count, k, mu_new, sd_new, px));
            MBRL_CUDA(cudaGetLastError());
          }
        }
        continue;
      } else {
        rc = launch_topk(p->d_costs, 1, p->N, kl, p->d_elite, p->d_ecost, nullptr, nullptr, it, st);
        if (rc) return rc;
        pack_elites_kernel<<<(kl + 255) / 256, 256, 0, st>>>(p->d_ecost, p->d_elite, kl, (int)cand_offset, p->d_send);
        MBRL_NCCL(g_nccl.AllGather(p->d_send, p->d_recv, (size_t)2 * kl, kNcclUint32, p->comm, st));
        unpack_gathered_kernel<<<(ng + 255) / 256, 256, 0, st>>>(p->d_recv, p->world, kl, p->d_gcost, p->d_gidx);
        rc = launch_topk(p->d_gcost, 1, ng, k, p->d_pos, nullptr, p->d_best_now, nullptr, it, st);
        if (rc) return rc;
        MBRL_CUDA(launch_pdl(remap_elites_kernel, dim3((std::max(k, p->world) + 255) / 256), dim3(256), 0, st, p->d_pos,
                             p->d_gidx, k, p->d_elite, p->d_best_now, p->d_best_ever, it, p->world, kl, k_full, p->d_trunc));
      }
      MBRL_CUDA(cudaGetLastError());
      if (it + 1 < I || need_final_dist) {
        // every rank holds the whole global list: the same per-rank partial sums, added in rank order
        ActionSource gsrc = src;
        gsrc.cand_offset = 0;
        rc = p->world == 1 ? launch_refit(p, gsrc, p->d_elite, k, p->d_mu_hist + (it + 1) * EHA, p->d_sd_hist + (it + 1) * EHA, st)
                           : launch_refit_seg(p, gsrc, p->d_elite, k, p->world, p->N, p->d_mu_hist + (it + 1) * EHA,
                                              p->d_sd_hist + (it + 1) * EHA, st);
        if (rc) return rc;
      }
      continue;
    }
    rc = launch_topk(p->d_costs, p->E, p->N, k, p->d_elite, nullptr, nullptr, p->d_best_ever, it, st, p->cur_stamps);
    if (rc) return rc;
    if (it + 1 < I || need_final_dist) {
      rc = p->refit_segments > 1 ? launch_refit_seg(p, src, p->d_elite, k, p->refit_segments, p->N / p->refit_segments,
                                                    p->d_mu_hist + (it + 1) * EHA, p->d_sd_hist + (it + 1) * EHA, st)
                                 : launch_refit(p, src, p->d_elite, k, p->d_mu_hist + (it + 1) * EHA, p->d_sd_hist + (it + 1) * EHA, st);
      if (rc) return rc;
    }
  }
  p->cur_stamps = nullptr;
  p->tc.fg.stamps = nullptr;
  int rc = launch_replay(p, a->sample_mode, a->seed, sharded ? 0u : a->cand_offset, a->env_offset, d_s0, d_injected,
                         p->d_mu_hist, p->d_sd_hist, I, a->return_mean, a->actions_only, p->d_best_ever, d_out_states,
                         d_out_actions, d_info, st);
  if (rc) return rc;
  if (warm_keep) {
    MBRL_CUDA(cudaMemcpyAsync(p->d_mu_last, p->d_mu_hist + (size_t)I * EHA, sizeof(float) * EHA, cudaMemcpyDeviceToDevice, st));
    p->have_last = true;
  } else {
    p->have_last = false;  // episode start (or a caller that does not warm start): forget the stored mean
  }
  if (sharded) {
    // info.reserved bit 0: the reduced-size elite gather was not provably exact; bit 1: the peer-memory
    // exchange timed out.  The truncation flag is reset here even when the caller passed no info buffer.
    MBRL_CUDA(launch_pdl(shard_flags_kernel, dim3(1), dim3(32), 0, st, d_info, p->d_trunc,
                         p->p2p_attached ? (const int*)p->d_p2p_error : (const int*)nullptr));
  }
  return MBRL_OK;
}

extern "C" int mbrl_plan_gd(MbrlPlanner* p, const MbrlGdArgs* a, const float* h_s0, const float* h_init_actions,
                            float* h_out_states, float* h_out_actions, float* h_out_cost, int32_t* h_out_iters) {
  if (!p || !a) return fail(MBRL_E_INVALID, "null planner/args");
  MBRL_REQUIRE(h_s0 && h_init_actions && h_out_states && h_out_actions, "null host buffer");
  MBRL_REQUIRE(a->restarts >= 1 && a->restarts <= 65535, "restarts out of range [1, 65535]");
  MBRL_REQUIRE(a->iterations >= 1 && a->iterations <= 100000, "iterations out of range");
  MBRL_REQUIRE(a->lr > 0.f && a->beta1 >= 0.f && a->beta1 < 1.f && a->beta2 >= 0.f && a->beta2 < 1.f && a->eps > 0.f, "bad Adam parameters");
  MBRL_REQUIRE(p->E == 1, "the gradient planner plans for one environment (num_envs == 1)");
  if (!p->have_weights || !p->have_cost) return fail(MBRL_E_STATE, "mbrl_set_weights and mbrl_set_cost must be called before planning");
  if (p->cost_kind != MBRL_COST_SMOOTHABS_COSH)
    return fail(MBRL_E_UNSUPPORTED, "the gradient planner differentiates the SmoothAbs + Cosh cost only");
  MBRL_CUDA(cudaSetDevice(p->cfg.device));
  const GdLayout L = gd_layout(p->O, p->A, p->U, p->H);
  size_t smem = sizeof(float) * (size_t)L.total;
  if (smem > p->max_smem)
    return fail(MBRL_E_UNSUPPORTED, "horizon x hidden too large for the gradient planner: it keeps every step's activations "
                                    "(H*(O+A+2*hidden) floats) in shared memory");
  const int B = a->restarts, HA = p->H * p->A, HO1 = (p->H + 1) * p->O;
  float *d_s0 = nullptr, *d_init = nullptr, *d_st = nullptr, *d_ac = nullptr, *d_cost = nullptr;
  int* d_it = nullptr;
  auto cleanup = [&]() { for (void* q : {(void*)d_s0, (void*)d_init, (void*)d_st, (void*)d_ac, (void*)d_cost, (void*)d_it}) if (q) cudaFree(q); };
  cudaError_t e = cudaSuccess;
  auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; return r == cudaSuccess; };
  ok(dev_alloc(&d_s0, (size_t)p->O)); ok(dev_alloc(&d_init, (size_t)B * HA)); ok(dev_alloc(&d_st, (size_t)B * HO1));
  ok(dev_alloc(&d_ac, (size_t)B * HA)); ok(dev_alloc(&d_cost, (size_t)B)); ok(dev_alloc(&d_it, (size_t)B));
  cudaStream_t st = p->stream;
  if (e == cudaSuccess) {
    ok(cudaMemcpyAsync(d_s0, h_s0, sizeof(float) * p->O, cudaMemcpyHostToDevice, st));
    ok(cudaMemcpyAsync(d_init, h_init_actions, sizeof(float) * (size_t)B * HA, cudaMemcpyHostToDevice, st));
  }
  // W2 rides in shared memory when it fits next to the activations (hidden 200, H = 30: to the byte, almost)
  const bool w2_smem = smem + sizeof(float) * (size_t)p->U * p->U <= p->max_smem;
  if (w2_smem) smem += sizeof(float) * (size_t)p->U * p->U;
  const bool regw = gd_small_in_regs(p->O, p->A, p->U);  // W1 / W3 in registers
  auto kern = w2_smem ? (regw ? gd_plan_kernel<true, true> : gd_plan_kernel<true, false>)
                      : (regw ? gd_plan_kernel<false, true> : gd_plan_kernel<false, false>);
  if (e == cudaSuccess) ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (e == cudaSuccess) {
    GdParams gp{p->H, a->iterations, a->lr, a->stop_condition, a->beta1, a->beta2, a->eps};
    kern<<<B, kGdThreads, smem, st>>>(model_view(p), gp, d_s0, d_init, d_st, d_ac, d_cost, d_it);
    ok(cudaGetLastError());
    ok(cudaMemcpyAsync(h_out_states, d_st, sizeof(float) * (size_t)B * HO1, cudaMemcpyDeviceToHost, st));
    ok(cudaMemcpyAsync(h_out_actions, d_ac, sizeof(float) * (size_t)B * HA, cudaMemcpyDeviceToHost, st));
    if (h_out_cost) ok(cudaMemcpyAsync(h_out_cost, d_cost, sizeof(float) * B, cudaMemcpyDeviceToHost, st));
    if (h_out_iters) ok(cudaMemcpyAsync(h_out_iters, d_it, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    ok(cudaStreamSynchronize(st));
  }
  cleanup();
  if (e != cudaSuccess) return fail(MBRL_E_CUDA, std::string("mbrl_plan_gd: ") + cudaGetErrorString(e));
  return MBRL_OK;
}

extern "C" int mbrl_plan_device(MbrlPlanner* p, const MbrlPlanArgs* args, const float* d_s0,
                                const float* d_injected, float* d_out_states, float* d_out_actions,
                                MbrlPlanInfo* d_info, void* stream) {
  if (!p || !args) return fail(MBRL_E_INVALID, "null planner/args");
  MBRL_REQUIRE(d_s0 && d_out_states && d_out_actions, "null device buffer");
  MBRL_CUDA(cudaSetDevice(p->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;  // NULL = the CUDA default stream, as everywhere in CUDA
  return enqueue_plan(p, args, d_s0, d_injected, d_out_states, d_out_actions, d_info,
                      args->return_mean != 0, st);
}

extern "C" int mbrl_plan(MbrlPlanner* p, const MbrlPlanArgs* args, const float* h_s0, float* h_out_states,
                         float* h_out_actions, MbrlPlanInfo* h_info, float* h_out_mu, float* h_out_sd) {
  if (!p || !args) return fail(MBRL_E_INVALID, "null planner/args");
  MBRL_REQUIRE(h_s0 && h_out_states && h_out_actions, "null host buffer");
  MBRL_REQUIRE((h_out_mu == nullptr) == (h_out_sd == nullptr), "out_mu and out_sd must be given together");
  MBRL_CUDA(cudaSetDevice(p->cfg.device));
  cudaStream_t st = p->stream;
  const int O = p->O, A = p->A, H = p->H, E = p->E;
  const size_t EHA = (size_t)E * H * A;
  const float* d_inj = nullptr;
  const bool inject = args->sample_mode == MBRL_SAMPLE_INJECT_ACTIONS || args->sample_mode == MBRL_SAMPLE_INJECT_NOISE;
  if (inject) {
    MBRL_REQUIRE(args->h_injected, "injected sample mode without h_injected");
    MBRL_REQUIRE(args->iterations >= 1, "iterations must be >= 1");
    const size_t need = (size_t)args->iterations * H * (size_t)p->R * A;
    if (need > p->injected_cap) {
      if (p->d_injected) MBRL_CUDA(cudaFree(p->d_injected));
      p->d_injected = nullptr; p->injected_cap = 0;
      MBRL_CUDA(dev_alloc(&p->d_injected, need));
      p->injected_cap = need;
    }
    MBRL_CUDA(cudaMemcpyAsync(p->d_injected, args->h_injected, sizeof(float) * need, cudaMemcpyHostToDevice, st));
    d_inj = p->d_injected;
  }
  std::memcpy(p->h_s0, h_s0, sizeof(float) * E * O);
  const bool need_dist = h_out_mu != nullptr || args->return_mean != 0;
  // The plan (a few KB) goes straight into the handle's pinned, device-mapped host buffers: the last kernels
  // store it there with posted PCIe writes, so no copy engine has to be started after them (three
  // device-to-host copies cost ~10 us of latency at the end of a 0.5 ms plan).  MBRL_NO_ZERO_COPY=1: copies.
  static const bool zero_copy_ok = getenv("MBRL_NO_ZERO_COPY") == nullptr;
  float *m_states = nullptr, *m_actions = nullptr;
  MbrlPlanInfo* m_info = nullptr;
  const bool small = sizeof(float) * ((size_t)E * H * O + EHA) <= (256u << 10);  // big batches: the copy engine's bursts win
  const bool zc = zero_copy_ok && small && cudaHostGetDevicePointer((void**)&m_states, p->h_out_states, 0) == cudaSuccess &&
                  cudaHostGetDevicePointer((void**)&m_actions, p->h_out_actions, 0) == cudaSuccess &&
                  cudaHostGetDevicePointer((void**)&m_info, p->h_info, 0) == cudaSuccess;
  if (!zc) cudaGetLastError();
  // ... and the initial states come the same way: the plan's first kernel reads them from the pinned buffer
  float* m_s0 = nullptr;
  const bool zs = zero_copy_ok && sizeof(float) * (size_t)E * O <= (64u << 10) &&
                  cudaHostGetDevicePointer((void**)&m_s0, p->h_s0, 0) == cudaSuccess;
  if (!zs) {
    cudaGetLastError();
    MBRL_CUDA(cudaMemcpyAsync(p->d_s0, p->h_s0, sizeof(float) * E * O, cudaMemcpyHostToDevice, st));
  }
  p->stage_s0 = zs ? StageS0{m_s0, p->d_s0, E * O} : StageS0{nullptr, nullptr, 0};
  struct Unstage { MbrlPlanner* q; ~Unstage() { q->stage_s0 = StageS0{nullptr, nullptr, 0}; } } unstage{p};
  int rc = enqueue_plan(p, args, p->d_s0, d_inj, zc ? m_states : p->d_out_states, zc ? m_actions : p->d_out_actions,
                        zc ? m_info : p->d_info, need_dist, st);
  if (rc) return rc;
  if (!zc) {
    MBRL_CUDA(cudaMemcpyAsync(p->h_out_actions, p->d_out_actions, sizeof(float) * EHA, cudaMemcpyDeviceToHost, st));
    MBRL_CUDA(cudaMemcpyAsync(p->h_out_states, p->d_out_states, sizeof(float) * E * H * O, cudaMemcpyDeviceToHost, st));
    MBRL_CUDA(cudaMemcpyAsync(p->h_info, p->d_info, sizeof(MbrlPlanInfo) * E, cudaMemcpyDeviceToHost, st));
  }
  if (h_out_mu) {
    MBRL_CUDA(cudaMemcpyAsync(p->h_mu, p->d_mu_hist + (size_t)args->iterations * EHA, sizeof(float) * EHA, cudaMemcpyDeviceToHost, st));
    MBRL_CUDA(cudaMemcpyAsync(p->h_sd, p->d_sd_hist + (size_t)args->iterations * EHA, sizeof(float) * EHA, cudaMemcpyDeviceToHost, st));
  }
  MBRL_CUDA(cudaStreamSynchronize(st));
  if (p->p2p_attached && (p->h_info[0].reserved & 2))
    return fail(MBRL_E_CUDA, "peer-memory elite exchange timed out waiting for another rank (MBRL_P2P_TIMEOUT_S); the ranks' "
                             "sequence numbers are out of step now: mbrl_p2p_detach, then export / attach again (or mbrl_comm_init)");
  if ((p->comm || p->p2p_attached) && (p->h_info[0].reserved & 1) && !p->full_gather) {
    // Practically unreachable (the shards are i.i.d.): some rank's reduced elite list was used up.
    // The flag derives from the gathered data, so every rank sees it and redoes the plan in lockstep.
    p->full_gather = true;
    const int rc2 = mbrl_plan(p, args, h_s0, h_out_states, h_out_actions, h_info, h_out_mu, h_out_sd);
    p->full_gather = false;  // the redo succeeded (or failed for another reason): back to reduced-size gathers
    return rc2;
  }
  std::memcpy(h_out_actions, p->h_out_actions, sizeof(float) * EHA);
  std::memcpy(h_out_states, p->h_out_states, sizeof(float) * E * H * O);
  if (h_info) std::memcpy(h_info, p->h_info, sizeof(MbrlPlanInfo) * E);
  if (h_out_mu) {
    std::memcpy(h_out_mu, p->h_mu, sizeof(float) * EHA);
    std::memcpy(h_out_sd, p->h_sd, sizeof(float) * EHA);
  }
  return MBRL_OK;
}
