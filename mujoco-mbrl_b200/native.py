"""ctypes binding of the C ABI (include/mbrl_b200.h) and a thin object wrapper.

PyTorch is used here only for device memory and streams (``tensor.data_ptr()``,
``torch.cuda.current_stream()``); every computation is a call into libmbrl_b200.so.
There is no fallback: if the library cannot be loaded this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmbrl_b200.so")

ENGINE_SIMT_FP32, ENGINE_TC_BF16, ENGINE_TC_FP16 = 0, 1, 2
ENGINES = {"fp32": ENGINE_SIMT_FP32, "simt": ENGINE_SIMT_FP32, "bf16": ENGINE_TC_BF16, "fp16": ENGINE_TC_FP16}
SAMPLE_INJECT_ACTIONS, SAMPLE_INJECT_NOISE, SAMPLE_GAUSSIAN, SAMPLE_UNIFORM = 0, 1, 2, 3
COST_SMOOTHABS_COSH, COST_DMC_CARTPOLE_SWINGUP, COST_REWARD_HEAD, COST_DMC_HUMANOID_RUN = 0, 1, 2, 3
COST_DMC_CHEETAH_RUN, COST_DMC_WALKER_WALK = 4, 5
WARM_USE, WARM_KEEP = 1, 2

# every symbol include/mbrl_b200.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "mbrl_abi_version", "mbrl_last_error", "mbrl_create", "mbrl_destroy", "mbrl_set_weights",
    "mbrl_set_norm", "mbrl_set_cost", "mbrl_set_action_bounds", "mbrl_set_reward_head", "mbrl_plan", "mbrl_plan_device",
    "mbrl_rollout", "mbrl_sample", "mbrl_philox_raw", "mbrl_topk", "mbrl_refit", "mbrl_emit",
    "mbrl_tc_debug", "mbrl_nccl_unique_id", "mbrl_comm_init", "mbrl_comm_destroy",
    "mbrl_p2p_export", "mbrl_p2p_attach", "mbrl_p2p_detach", "mbrl_plan_gd", "mbrl_set_refit_segments",
]


class MbrlConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "obs_dim", "act_dim", "hidden", "horizon", "num_candidates", "num_envs", "max_iterations",
        "max_elites", "engine", "device")]


class MbrlPlanArgs(C.Structure):
    _fields_ = [
        ("iterations", C.c_int32), ("elites", C.c_int32), ("sample_mode", C.c_int32), ("return_mean", C.c_int32),
        ("seed", C.c_uint64), ("cand_offset", C.c_uint32), ("env_offset", C.c_uint32),
        ("h_injected", C.c_void_p), ("h_mu0", C.c_void_p), ("h_sd0", C.c_void_p),
        ("actions_only", C.c_int32), ("warm_start", C.c_int32), ("warm_std", C.c_float), ("reserved", C.c_int32),
    ]


class MbrlGdArgs(C.Structure):
    _fields_ = [("restarts", C.c_int32), ("iterations", C.c_int32), ("lr", C.c_float), ("stop_condition", C.c_float),
                ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("reserved", C.c_int32)]


class MbrlPlanInfo(C.Structure):
    _fields_ = [("best_cost", C.c_float), ("best_iteration", C.c_int32), ("best_index", C.c_int32),
                ("reserved", C.c_int32)]


PLAN_INFO_DTYPE = np.dtype([("best_cost", "<f4"), ("best_iteration", "<i4"), ("best_index", "<i4"), ("reserved", "<i4")])


class MbrlError(RuntimeError):
    pass


_lib = None


def load_library(build_if_missing: bool = True) -> C.CDLL:
    """Load libmbrl_b200.so (building it with nvcc when missing/stale and a compiler is
    present).  Raises -- never falls back -- when the library is unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        try:
            from . import build as _build
            if _build.needs_build():
                _build.build()
        except Exception as exc:  # stale-but-present libraries are still usable
            if not os.path.exists(LIB_PATH):
                raise MbrlError(f"libmbrl_b200.so is missing and could not be built: {exc}") from exc
    if not os.path.exists(LIB_PATH):
        raise MbrlError(f"{LIB_PATH} not found; run `python mujoco-mbrl_b200/build.py` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    p, i32, u32, u64, i64, f32, f64, vp = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_int64, C.c_float, C.c_double, C.c_void_p
    lib.mbrl_abi_version.restype = C.c_int
    lib.mbrl_last_error.restype = C.c_char_p
    sigs = {
        "mbrl_create": [C.POINTER(MbrlConfig), C.POINTER(p)],
        "mbrl_destroy": [p],
        "mbrl_set_weights": [p] + [vp] * 6,
        "mbrl_set_norm": [p] + [vp] * 4,
        "mbrl_set_cost": [p, i32, vp, vp, f64, f64],
        "mbrl_set_action_bounds": [p, f32, f32],
        "mbrl_set_reward_head": [p, vp, f32, f32, f32],
        "mbrl_plan": [p, C.POINTER(MbrlPlanArgs), vp, vp, vp, vp, vp, vp],
        "mbrl_plan_device": [p, C.POINTER(MbrlPlanArgs), vp, vp, vp, vp, vp, vp],
        "mbrl_plan_gd": [p, C.POINTER(MbrlGdArgs), vp, vp, vp, vp, vp, vp],
        "mbrl_rollout": [p, i32, u64, u32, u32, u32, vp, vp, vp, vp, vp, vp, vp, vp],
        "mbrl_sample": [p, i32, u64, u32, u32, u32, vp, vp, vp, vp],
        "mbrl_philox_raw": [vp, vp, vp, i64, vp],
        "mbrl_topk": [vp, i32, i32, i32, vp, vp, vp, vp],
        "mbrl_refit": [p, i32, u64, u32, u32, u32, vp, vp, vp, vp, i32, vp, vp, vp],
        "mbrl_emit": [p, i32, u64, u32, u32, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp],
        "mbrl_tc_debug": [p, i32, vp],
        "mbrl_nccl_unique_id": [vp],
        "mbrl_comm_init": [p, vp, i32, i32],
        "mbrl_comm_destroy": [p],
        "mbrl_p2p_export": [p, i32, vp],
        "mbrl_p2p_attach": [p, vp, i32, i32],
        "mbrl_p2p_detach": [p],
        "mbrl_set_refit_segments": [p, i32],
    }
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        msg = load_library().mbrl_last_error()
        raise MbrlError(f"mbrl_b200 error {rc}: {msg.decode() if msg else '?'}")


def _f32(x) -> np.ndarray:
    """Host fp32 C-contiguous view/copy of a tensor/array."""
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=np.float32)


def _hp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _dp(t):
    """Device pointer of a CUDA torch tensor (None passes NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return C.c_void_p(t.data_ptr())


def _stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class NativePlanner:
    """One C-ABI handle: fixed (O, A, U, H, N, E) on one device."""

    def __init__(self, obs_dim, act_dim, hidden, horizon, num_candidates, num_envs=1, max_iterations=1,
                 max_elites=None, engine="fp32", device=0):
        self.lib = load_library()
        self.O, self.A, self.U, self.H, self.N, self.E = obs_dim, act_dim, hidden, horizon, num_candidates, num_envs
        self.R = num_candidates * num_envs
        self.max_iterations = max_iterations
        self.max_elites = num_candidates if max_elites is None else max_elites
        self.engine = ENGINES[engine] if isinstance(engine, str) else int(engine)
        self.device = device
        cfg = MbrlConfig(obs_dim, act_dim, hidden, horizon, num_candidates, num_envs, max_iterations,
                         self.max_elites, self.engine, device)
        handle = C.c_void_p()
        _check(self.lib.mbrl_create(C.byref(cfg), C.byref(handle)))
        self._h = handle

    def close(self):
        if getattr(self, "_h", None):
            self.lib.mbrl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- model / cost upload ---------------------------------------------------------
    def set_weights(self, W1, b1, W2, b2, W3, b3):
        arrs = [_f32(a) for a in (W1, b1, W2, b2, W3, b3)]
        shapes = [(self.U, self.O + self.A), (self.U,), (self.U, self.U), (self.U,), (self.O, self.U), (self.O,)]
        for a, s in zip(arrs, shapes):
            if a.shape != s:
                raise ValueError(f"weight shape {a.shape} != expected {s}")
        _check(self.lib.mbrl_set_weights(self._h, *[_hp(a) for a in arrs]))

    def set_norm(self, mu_s=None, sd_s=None, mu_a=None, sd_a=None):
        arrs = [None if a is None else _f32(a).reshape(-1) for a in (mu_s, sd_s, mu_a, sd_a)]
        for a, n in zip(arrs, (self.O, self.O, self.A, self.A)):
            if a is not None and a.shape != (n,):
                raise ValueError(f"stat shape {a.shape} != ({n},)")
        _check(self.lib.mbrl_set_norm(self._h, *[_hp(a) for a in arrs]))

    def set_cost(self, weights=None, goal=None, alpha=0.4, beta=0.25, kind=COST_SMOOTHABS_COSH):
        if kind != COST_SMOOTHABS_COSH:
            _check(self.lib.mbrl_set_cost(self._h, kind, None, None, float(alpha), float(beta)))
            return
        w, g = _f32(weights).reshape(-1), _f32(goal).reshape(-1)
        if w.shape != (self.O,) or g.shape != (self.O,):
            raise ValueError("cost weights/goal must have obs_dim entries")
        _check(self.lib.mbrl_set_cost(self._h, kind, _hp(w), _hp(g), float(alpha), float(beta)))

    def set_reward_head(self, W4, b4, reward_mean=0.0, reward_std=1.0):
        """ModelWithReward.linear4 + the "rewards" statistics (src/mbrl/models.py:132, agents.py:347)."""
        w = _f32(W4).reshape(-1)
        if w.shape != (self.U,):
            raise ValueError(f"reward head must have {self.U} weights, got {w.shape}")
        scalar = lambda v: float(_f32(v).reshape(-1)[0])
        _check(self.lib.mbrl_set_reward_head(self._h, _hp(w), scalar(b4), scalar(reward_mean), scalar(reward_std)))

    def set_action_bounds(self, lo, hi):
        _check(self.lib.mbrl_set_action_bounds(self._h, float(lo), float(hi)))

    def load_problem(self, prob):
        """Upload a PlanningProblem (adaptor.py)."""
        self.set_weights(prob.W1, prob.b1, prob.W2, prob.b2, prob.W3, prob.b3)
        self.set_norm(prob.mu_s, prob.sd_s, prob.mu_a, prob.sd_a)
        if getattr(prob, "W4", None) is not None:
            self.set_reward_head(prob.W4, prob.b4, prob.mu_r, prob.sd_r)
        self.set_cost(prob.cost_w, prob.goal, prob.alpha, prob.beta, getattr(prob, "cost_kind", COST_SMOOTHABS_COSH))
        self.set_action_bounds(prob.act_lo, prob.act_hi)

    # ---- whole plans -----------------------------------------------------------------
    def _args(self, iterations, elites, mode, seed, return_mean, cand_offset, env_offset, injected, mu0, sd0,
              actions_only=False, warm_start=0, warm_std=0.0):
        keep = []
        a = MbrlPlanArgs()
        a.actions_only = int(actions_only)
        a.warm_start, a.warm_std = int(warm_start), float(warm_std)
        a.iterations, a.elites, a.sample_mode, a.return_mean = iterations, elites, mode, int(return_mean)
        a.seed, a.cand_offset, a.env_offset = int(seed) & (2 ** 64 - 1), cand_offset, env_offset
        for name, val, n in (("h_injected", injected, iterations * self.H * self.R * self.A),
                             ("h_mu0", mu0, self.E * self.H * self.A), ("h_sd0", sd0, self.E * self.H * self.A)):
            if val is not None:
                arr = _f32(val).reshape(-1)
                if arr.size != n:
                    raise ValueError(f"{name}: {arr.size} elements, expected {n}")
                keep.append(arr)
                setattr(a, name, arr.ctypes.data)
        return a, keep

    def plan(self, s0, iterations=1, elites=1, mode=SAMPLE_GAUSSIAN, seed=0, injected=None, mu0=None, sd0=None,
             return_mean=False, want_dist=False, cand_offset=0, env_offset=0, actions_only=False, warm_start=0,
             warm_std=0.0):
        """mbrl_plan with host buffers.  s0: [O] or [E,O].  Returns a dict of numpy arrays:
        states [E,H,O], actions [E,H,A], info (structured [E]), optionally mu/sd [E,H,A]."""
        s0 = _f32(s0).reshape(self.E, self.O)
        args, keep = self._args(iterations, elites, mode, seed, return_mean, cand_offset, env_offset, injected, mu0, sd0,
                                actions_only, warm_start, warm_std)
        states = np.empty((self.E, self.H, self.O), np.float32)
        actions = np.empty((self.E, self.H, self.A), np.float32)
        info = np.zeros(self.E, PLAN_INFO_DTYPE)
        mu = np.empty((self.E, self.H, self.A), np.float32) if want_dist else None
        sd = np.empty((self.E, self.H, self.A), np.float32) if want_dist else None
        _check(self.lib.mbrl_plan(self._h, C.byref(args), _hp(s0), _hp(states), _hp(actions),
                                  info.ctypes.data_as(C.c_void_p), _hp(mu), _hp(sd)))
        del keep
        return dict(states=states, actions=actions, info=info, mu=mu, sd=sd)

    def plan_device(self, d_s0, d_out_states, d_out_actions, d_info=None, iterations=1, elites=1,
                    mode=SAMPLE_GAUSSIAN, seed=0, d_injected=None, return_mean=False, cand_offset=0, env_offset=0,
                    actions_only=False, warm_start=0, warm_std=0.0):
        """mbrl_plan_device: everything resident in HBM, enqueued on torch's current stream."""
        args, keep = self._args(iterations, elites, mode, seed, return_mean, cand_offset, env_offset, None, None, None,
                                actions_only, warm_start, warm_std)
        _check(self.lib.mbrl_plan_device(self._h, C.byref(args), _dp(d_s0), _dp(d_injected), _dp(d_out_states),
                                         _dp(d_out_actions), _dp(d_info), _stream_ptr()))

    def plan_gd(self, s0, init_actions, iterations=40, stop_condition=0.002, lr=0.01, betas=(0.9, 0.999), eps=1e-8):
        """mbrl_plan_gd: the batched GradientDescentPlanner.  init_actions [B,H,A] (or [H,A]).  Returns a dict of
        numpy arrays: states [B,H+1,O], actions [B,H,A], cost [B], iterations [B]."""
        s0 = _f32(s0).reshape(self.O)
        init = _f32(init_actions).reshape(-1, self.H, self.A)
        B = init.shape[0]
        args = MbrlGdArgs(B, int(iterations), float(lr), float(stop_condition), float(betas[0]), float(betas[1]), float(eps), 0)
        states = np.empty((B, self.H + 1, self.O), np.float32)
        actions = np.empty((B, self.H, self.A), np.float32)
        cost = np.empty(B, np.float32)
        iters = np.empty(B, np.int32)
        _check(self.lib.mbrl_plan_gd(self._h, C.byref(args), _hp(s0), _hp(init), _hp(states), _hp(actions), _hp(cost), _hp(iters)))
        return dict(states=states, actions=actions, cost=cost, iterations=iters)

    # ---- building blocks (device tensors) --------------------------------------------
    def rollout(self, d_s0, mode, seed=0, iteration=0, d_injected=None, d_mu=None, d_sd=None, want_states=False,
                want_actions=False, cand_offset=0, env_offset=0):
        import torch
        dev = d_s0.device
        costs = torch.empty(self.R, dtype=torch.float32, device=dev)
        states = torch.empty(self.H * self.R, self.O, dtype=torch.float32, device=dev) if want_states else None
        actions = torch.empty(self.H * self.R, self.A, dtype=torch.float32, device=dev) if want_actions else None
        _check(self.lib.mbrl_rollout(self._h, mode, int(seed), iteration, cand_offset, env_offset, _dp(d_s0),
                                     _dp(d_injected), _dp(d_mu), _dp(d_sd), _dp(costs), _dp(states), _dp(actions),
                                     _stream_ptr()))
        return costs, states, actions

    def sample(self, mode, seed=0, iteration=0, d_mu=None, d_sd=None, cand_offset=0, env_offset=0, device=None):
        import torch
        dev = device if device is not None else (d_mu.device if d_mu is not None else f"cuda:{self.device}")
        out = torch.empty(self.H * self.R, self.A, dtype=torch.float32, device=dev)
        _check(self.lib.mbrl_sample(self._h, mode, int(seed), iteration, cand_offset, env_offset, _dp(d_mu), _dp(d_sd),
                                    _dp(out), _stream_ptr()))
        return out

    def refit(self, d_elite_idx, k, mode, seed=0, iteration=0, d_injected=None, d_mu=None, d_sd=None,
              cand_offset=0, env_offset=0):
        import torch
        dev = d_elite_idx.device
        mu = torch.empty(self.E, self.H, self.A, dtype=torch.float32, device=dev)
        sd = torch.empty_like(mu)
        _check(self.lib.mbrl_refit(self._h, mode, int(seed), iteration, cand_offset, env_offset, _dp(d_injected),
                                   _dp(d_mu), _dp(d_sd), _dp(d_elite_idx), k, _dp(mu), _dp(sd), _stream_ptr()))
        return mu, sd


    def comm_init(self, rank=None, world=None, group=None):
        """Make this handle one shard of a population split over the ranks of a torch.distributed
        group (NCCL inside the library; rank 0's unique id is broadcast through torch)."""
        import torch
        import torch.distributed as dist
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        buf = np.zeros(128, np.uint8)
        if rank == 0:
            _check(self.lib.mbrl_nccl_unique_id(_hp(buf)))
        if world > 1:
            t = torch.from_numpy(buf)
            if dist.get_backend(group) == "nccl":
                t = t.cuda(self.device)
            dist.broadcast(t, src=0, group=group)
            buf = t.cpu().numpy().copy()
        _check(self.lib.mbrl_comm_init(self._h, _hp(buf), rank, world))
        self.rank, self.world = rank, world

    def p2p_init(self, rank=None, world=None, group=None):
        """Peer-memory (NVLink P2P, CUDA IPC) transport for the sharded loop: export this rank's gather
        buffer, all-gather the 64-byte IPC handles through torch.distributed, open the peers' buffers.
        Collective and all-or-nothing: every rank takes part in both exchanges whatever happened
        locally, and if ANY rank failed to export or attach, every rank detaches and the call returns
        False (the caller then uses comm_init: a mix of transports would stall the elite exchange)."""
        import torch
        import torch.distributed as dist
        if world is None:
            world, rank = dist.get_world_size(group), dist.get_rank(group)
        on_gpu = dist.get_backend(group) == "nccl"
        dev = torch.device("cuda", self.device) if on_gpu else torch.device("cpu")
        mine = np.zeros(64, np.uint8)
        ok = self.lib.mbrl_p2p_export(self._h, world, _hp(mine)) == 0
        allh = torch.empty(world * 64, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, torch.from_numpy(mine).to(dev), group=group)
        if ok:
            handles = np.ascontiguousarray(allh.cpu().numpy())
            ok = self.lib.mbrl_p2p_attach(self._h, _hp(handles), rank, world) == 0
        err = None if ok else self.lib.mbrl_last_error().decode(errors="replace")
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            self.lib.mbrl_p2p_detach(self._h)
            self.p2p_error = err
            return False
        self.rank, self.world = rank, world
        return True

    def set_refit_segments(self, segments):
        """mbrl_set_refit_segments: make this UNSHARDED planner add the refit's sums the way a plan
        sharded over `segments` ranks does (per-rank partial sums in rank order) -- bit-identical to it."""
        _check(self.lib.mbrl_set_refit_segments(self._h, int(segments)))

    def tc_debug(self, enable=True, fetch=False):
        """Diagnostic: arm / fetch the raw accumulator dump of tile 0, step 0 ([3,128,512]: three
        layers, 128 rows, 512-column pitch); the clock64 timeline of tile 1 ([64,32] int64) is left
        in `self.tc_timeline`."""
        nb = 3 * 128 * 512 * 4
        buf = np.zeros(nb + 64 * 32 * 8, np.uint8) if fetch else None
        _check(self.lib.mbrl_tc_debug(self._h, int(enable), _hp(buf)))
        if buf is None:
            return None
        self.tc_timeline = buf[nb:].view(np.int64).reshape(64, 32).copy()
        return buf[:nb].view(np.float32).reshape(3, 128, 512).copy()

    def emit(self, d_s0, d_best, d_mu_hist, d_sd_hist, iterations, mode, seed=0, d_injected=None, return_mean=False,
             cand_offset=0, env_offset=0):
        """mbrl_emit: (states [E,H,O], actions [E,H,A]) of the candidates named by d_best
        ([E,4] int32: cost bits, iteration, index, 0)."""
        import torch
        dev = d_s0.device
        states = torch.empty(self.E, self.H, self.O, dtype=torch.float32, device=dev)
        actions = torch.empty(self.E, self.H, self.A, dtype=torch.float32, device=dev)
        _check(self.lib.mbrl_emit(self._h, mode, int(seed), cand_offset, env_offset, _dp(d_s0), _dp(d_injected),
                                  _dp(d_mu_hist), _dp(d_sd_hist), iterations, int(return_mean), _dp(d_best),
                                  _dp(states), _dp(actions), _stream_ptr()))
        return states, actions


def topk(d_costs, k, segments=1):
    """mbrl_topk on a CUDA float32 tensor of `segments` equal-length cost arrays.
    Returns (elite_idx [segments,k] int32 ascending index order, elite_cost, best (structured numpy))."""
    import torch
    lib = load_library()
    n = d_costs.numel() // segments
    idx = torch.empty(segments, k, dtype=torch.int32, device=d_costs.device)
    cost = torch.empty(segments, k, dtype=torch.float32, device=d_costs.device)
    best = torch.zeros(segments, 4, dtype=torch.int32, device=d_costs.device)
    _check(lib.mbrl_topk(_dp(d_costs), segments, n, k, _dp(idx), _dp(cost), _dp(best), _stream_ptr()))
    return idx, cost, best


def philox_raw(d_ctr, d_key):
    import torch
    lib = load_library()
    out = torch.empty_like(d_ctr)
    _check(lib.mbrl_philox_raw(_dp(d_ctr), _dp(d_key), _dp(out), d_ctr.shape[0], _stream_ptr()))
    return out
