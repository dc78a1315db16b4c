"""numpy restatement of this repo's counter-based action sampler (TEST INFRASTRUCTURE).

This is NOT reference behaviour: the reference draws fp64 uniforms from numpy's global
MT19937 (src/mbrl/env_wrappers.py:50-62), which a counter-based GPU generator cannot
reproduce by design.  Parity runs therefore inject recorded noise; this file pins the
generator itself:

  * Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11),
    checked in tests against the Random123 known-answer vectors;
  * the counter/key layout and the uint32 -> uniform -> Box-Muller mapping that
    ``mujoco-mbrl_b200/csrc/philox.cuh`` implements.

Counter layout (one Philox call yields the 4 draws for action dims 4g..4g+3 of one
candidate at one step):
    ctr = (h * G + g,  iteration,  global candidate index,  environment index)
    key = (seed & 0xffffffff, seed >> 32)            with G = ceil(A / 4)
so results are independent of how candidates / environments are sharded over GPUs.
"""
from __future__ import annotations

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr: [..., 4] uint32, key: [..., 2] uint32 (broadcastable) -> [..., 4] uint32."""
    c = [np.asarray(ctr[..., i], dtype=np.uint32) for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32)
    k1 = np.asarray(key[..., 1], dtype=np.uint32)
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = _M0 * c[0].astype(np.uint64)
            p1 = _M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            if r != 9:
                k0 = (k0 + _W0).astype(np.uint32)
                k1 = (k1 + _W1).astype(np.uint32)
    return np.stack(np.broadcast_arrays(*c), axis=-1)


def u32_to_uniform(x: np.ndarray) -> np.ndarray:
    """((x >> 9) + 0.5) * 2^-23: exact in fp32 (24 significant bits), strictly inside (0, 1)."""
    return ((x >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)


def box_muller(u: np.ndarray) -> np.ndarray:
    """[..., 4] uniforms -> [..., 4] standard normals:
    (r0 cos t0, r0 sin t0, r1 cos t1, r1 sin t1), r = sqrt(-2 ln u_even), t = 2 pi u_odd."""
    u = u.astype(np.float32)
    r0 = np.sqrt(np.float32(-2.0) * np.log(u[..., 0]))
    r1 = np.sqrt(np.float32(-2.0) * np.log(u[..., 2]))
    t0 = (np.float64(2.0) * np.pi * u[..., 1].astype(np.float64))
    t1 = (np.float64(2.0) * np.pi * u[..., 3].astype(np.float64))
    z = np.stack(
        [r0 * np.cos(t0).astype(np.float32), r0 * np.sin(t0).astype(np.float32),
         r1 * np.cos(t1).astype(np.float32), r1 * np.sin(t1).astype(np.float32)],
        axis=-1,
    )
    return z.astype(np.float32)


def _counters(seed: int, iteration: int, horizon: int, n: int, act_dim: int,
              cand_offset: int = 0, env: int = 0):
    G = (act_dim + 3) // 4
    h = np.arange(horizon, dtype=np.uint32)[:, None, None]
    c = (np.arange(n, dtype=np.uint32) + np.uint32(cand_offset))[None, :, None]
    g = np.arange(G, dtype=np.uint32)[None, None, :]
    ctr = np.stack(
        np.broadcast_arrays(h * np.uint32(G) + g, np.uint32(iteration), c, np.uint32(env)), axis=-1
    ).astype(np.uint32)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    return ctr, key, G


def raw_u32(seed, iteration, horizon, n, act_dim, cand_offset=0, env=0) -> np.ndarray:
    """[H, N, G*4] raw Philox words in action-dim order."""
    ctr, key, G = _counters(seed, iteration, horizon, n, act_dim, cand_offset, env)
    return philox4x32_10(ctr, key).reshape(horizon, n, G * 4)


def standard_normal(seed, iteration, horizon, n, act_dim, cand_offset=0, env=0) -> np.ndarray:
    """[H*N, A] step-major standard normals, as the device sampler defines them."""
    ctr, key, G = _counters(seed, iteration, horizon, n, act_dim, cand_offset, env)
    z = box_muller(u32_to_uniform(philox4x32_10(ctr, key)))  # [H, N, G, 4]
    return z.reshape(horizon, n, G * 4)[:, :, :act_dim].reshape(horizon * n, act_dim)


def uniform(seed, iteration, horizon, n, act_dim, lo, hi, cand_offset=0, env=0) -> np.ndarray:
    """[H*N, A] step-major uniforms lo + (hi-lo)*u (random-shooting mode; the reference
    sampler's distribution, src/mbrl/env_wrappers.py:52-62)."""
    ctr, key, G = _counters(seed, iteration, horizon, n, act_dim, cand_offset, env)
    u = u32_to_uniform(philox4x32_10(ctr, key)).reshape(horizon, n, G * 4)[:, :, :act_dim]
    out = np.float32(lo) + (np.float32(hi) - np.float32(lo)) * u
    return out.reshape(horizon * n, act_dim).astype(np.float32)
