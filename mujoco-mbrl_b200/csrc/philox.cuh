// Counter-based action sampler: Philox4x32-10 + Box-Muller + clip, usable from any kernel
// so that the rollout kernel and the refit kernel regenerate bit-identical actions and no
// action tensor has to round-trip HBM.
//
// Replaces EnvWrapper._sample_action (src/mbrl/env_wrappers.py:50-62: numpy MT19937
// uniforms, one call for all H*N rows) -- distribution-equivalent in MBRL_SAMPLE_UNIFORM
// mode, not bit-equivalent (parity runs inject recorded draws instead).
//
// Counter layout (mirrored by oracle/philox.py):
//   ctr = (h*G + g, iteration, global candidate, global env),  key = seed,  G = ceil(A/4)
// One call yields the draws for action dims 4g..4g+3 of one candidate at one step.
#pragma once
#include "common.cuh"

namespace mbrl {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    if (r != 9) {
      k.x += 0x9E3779B9u;
      k.y += 0xBB67AE85u;
    }
  }
  return c;
}

// ((x >> 9) + 0.5) * 2^-23: exact in fp32, strictly inside (0,1)
__device__ __forceinline__ float u32_to_uniform(uint32_t x) {
  return __fmul_rn(__fadd_rn((float)(x >> 9), 0.5f), 1.1920928955078125e-07f);
}

// Box-Muller on the SFU: lg2.approx, sqrt.approx, sin.approx / cos.approx (absolute error <= 2^-21.4 on
// [-pi, pi]; the angle 2*pi*u is folded there exactly: u - 0.5 is exact, sin(2 pi u) = -sin(2 pi (u - 0.5))).
// A draw differs from the libm formula (oracle/philox.py) by <= 4e-6 in absolute terms; every consumer
// -- rollout samplers, refit, replay, sample_kernel -- calls this one function, so they agree bit for bit.
// The libm version (logf, sincospif: ~130 more instructions per 4 draws) made the sampler warps of the
// fused rollout kernel its critical path: 92.2 us per cfg-3 rollout against 86.0 us with uniform draws.
__device__ __forceinline__ float4 box_muller4(uint4 r) {
  const float u0 = u32_to_uniform(r.x), u1 = u32_to_uniform(r.y);
  const float u2 = u32_to_uniform(r.z), u3 = u32_to_uniform(r.w);
  float r0, r1;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(__fmul_rn(-2.0f, __logf(u0))));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(__fmul_rn(-2.0f, __logf(u2))));
  const float x0 = __fmul_rn(6.283185307179586f, __fsub_rn(u1, 0.5f));
  const float x1 = __fmul_rn(6.283185307179586f, __fsub_rn(u3, 0.5f));
  const float s0 = -__sinf(x0), c0 = -__cosf(x0), s1 = -__sinf(x1), c1 = -__cosf(x1);
  return make_float4(__fmul_rn(r0, c0), __fmul_rn(r0, s0), __fmul_rn(r1, c1), __fmul_rn(r1, s1));
}

// Calls emit(a, value) for a = 0..A-1 with the action of (step h, local env, local cand).
// row = env_l*N + cand_l;  R = rows on this GPU.  mu/sd are indexed with the LOCAL env,
// the Philox counter with GLOBAL env / candidate indices (shard-independent streams).
template <class Emit>
__device__ __forceinline__ void for_each_action(const ActionSource& s, int A, int H, int h,
                                                int env_l, int cand_l, long long row,
                                                long long R, Emit&& emit) {
  const long long ms = ((long long)env_l * H + h) * A;
  if (s.mode == MBRL_SAMPLE_INJECT_ACTIONS) {
    const float* p = s.buf + ((long long)h * R + row) * A;
    for (int a = 0; a < A; ++a) emit(a, __ldg(p + a));
  } else if (s.mode == MBRL_SAMPLE_INJECT_NOISE) {
    const float* p = s.buf + ((long long)h * R + row) * A;
    for (int a = 0; a < A; ++a) {
      // clamp(mu + sd*z): separate mul and add, as the torch op chain rounds
      const float v = __fadd_rn(dep_load(s.mu + ms + a), __fmul_rn(dep_load(s.sd + ms + a), __ldg(p + a)));
      emit(a, clipf(v, s.lo, s.hi));
    }
  } else {
    const int G = (A + 3) >> 2;
    const uint2 key = make_uint2(s.seed_lo, s.seed_hi);
    for (int g = 0; g < G; ++g) {
      const uint4 r = philox4x32_10(
          make_uint4((uint32_t)(h * G + g), s.iteration, s.cand_offset + (uint32_t)cand_l,
                     s.env_offset + (uint32_t)env_l),
          key);
      float v[4];
      if (s.mode == MBRL_SAMPLE_GAUSSIAN) {
        const float4 z = box_muller4(r);
        v[0] = z.x; v[1] = z.y; v[2] = z.z; v[3] = z.w;
      } else {
        v[0] = u32_to_uniform(r.x); v[1] = u32_to_uniform(r.y);
        v[2] = u32_to_uniform(r.z); v[3] = u32_to_uniform(r.w);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int a = 4 * g + j;
        if (a < A) {
          float out;
          if (s.mode == MBRL_SAMPLE_GAUSSIAN)
            out = clipf(__fadd_rn(dep_load(s.mu + ms + a), __fmul_rn(dep_load(s.sd + ms + a), v[j])), s.lo, s.hi);
          else
            out = __fadd_rn(s.lo, __fmul_rn(__fsub_rn(s.hi, s.lo), v[j]));
          emit(a, out);
        }
      }
    }
  }
}

// ---- standalone kernels -------------------------------------------------------------

__global__ void philox_raw_kernel(const uint32_t* __restrict__ ctr, const uint32_t* __restrict__ key,
                                  uint32_t* __restrict__ out, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 c = make_uint4(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3]);
  const uint4 r = philox4x32_10(c, make_uint2(key[2 * i], key[2 * i + 1]));
  out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

// Materialises the sampler output, [H*R, A] step-major.  One thread per (h, row); HBM-bound
// (4*A bytes written per thread) -- used for tests and for the sampling-bandwidth figure.
__global__ void sample_kernel(ActionSource src, Shape sh, int A, float* __restrict__ out) {
  const long long R = sh.rows();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= R * sh.H) return;
  const int h = (int)(i / R);
  const long long row = i - (long long)h * R;
  const int env_l = (int)(row / sh.N), cand_l = (int)(row - (long long)env_l * sh.N);
  float* o = out + i * A;
  for_each_action(src, A, sh.H, h, env_l, cand_l, row, R, [&](int a, float v) { o[a] = v; });
}

}  // namespace mbrl
