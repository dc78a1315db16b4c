"""Drop-in planners with the reference's API (src/mbrl/planners.py:14-25, 140-187).

    states, actions = RandomShootingPlanner.plan(initial_state, model, cost, sample_action, horizon,
                                                 initial_trajectory=None, num_trajectories=1000)
    states, actions = CEMPlanner.plan(initial_state, model, cost, sample_action, horizon,
                                      initial_trajectory=None, num_trajectories=16384,
                                      num_iterations=5, elite_frac=0.1)

Planners are passed around as classes with static methods and pickled with the policy
(src/mbrl/experiment.py:22-26, src/mbrl/parallel.py:23-38), so no CUDA handle lives on them:
handles sit in a lazily built module-level cache keyed by model identity and shape.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from . import native
from .adaptor import PlanningProblem, problem_from_callables, problem_from_callables_cached

_HANDLES: Dict[tuple, dict] = {}
_HANDLES_PID = None  # the process that built the cache: a forked child must not reuse the parent's CUDA handles
_SEED_SALT = None    # per-process salt of the default seed stream (parallel rollout workers explore differently)


def _default_seed(ent):
    """seed=None: a per-handle call counter on top of a per-process random salt, so that the reference's
    parallel rollout workers (src/mbrl/parallel.py:20-38), which all start their counters at zero, do not
    evaluate identical Philox candidate sets at the same MPC step."""
    global _SEED_SALT
    if _SEED_SALT is None:
        import os
        _SEED_SALT = int.from_bytes(os.urandom(6), "little") << 16
    return _SEED_SALT + ent["calls"]


def _as_torch(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x))


def _check_process():
    """Handles hold CUDA state that does not survive fork(): each process builds its own cache, and a
    forked child of a parent that had already initialised CUDA gets a clear error instead of a cryptic
    CUDA initialisation failure inside mbrl_create (use the 'spawn' or 'forkserver' start method, as the
    reference's entry point does: src/mbrl/experiment.py:203)."""
    global _HANDLES_PID
    import os
    pid = os.getpid()
    if _HANDLES_PID is None:
        _HANDLES_PID = pid
    elif _HANDLES_PID != pid:
        _HANDLES.clear()  # the parent's handles are not ours to destroy: just forget them
        _HANDLES_PID = pid
        try:
            import torch
            bad_fork = getattr(torch.cuda, "_is_in_bad_fork", None)
            if (bad_fork() if bad_fork else torch.cuda.is_initialized()):
                raise native.MbrlError(
                    "this process was fork()ed from a parent that had already initialised CUDA; CUDA cannot be "
                    "used in such a child.  Start rollout workers with the 'spawn' or 'forkserver' method "
                    "(torch.multiprocessing.set_start_method), as src/mbrl/experiment.py:203 does")
        except ImportError:
            pass


def _get_handle(model, cost, sample_action, horizon, n, max_iters, engine, device):
    _check_process()
    prob, fp = problem_from_callables_cached(model, cost, sample_action)
    owner = model.func if hasattr(model, "func") else model
    key = (id(owner), prob.obs_dim, prob.act_dim, prob.hidden, horizon, n, max_iters, engine, device)
    ent = _HANDLES.get(key)
    if ent is None:
        h = native.NativePlanner(prob.obs_dim, prob.act_dim, prob.hidden, horizon, n, 1, max_iters, n, engine, device)
        ent = dict(handle=h, fingerprint=None, calls=0, keepalive=None)
        _HANDLES[key] = ent
    if ent["fingerprint"] != fp:
        ent["handle"].load_problem(prob)
        ent["fingerprint"] = fp
        ent["keepalive"] = (owner, prob)  # keeps ids in the fingerprint from being recycled
    return ent, prob


def clear_handles():
    """Destroy every cached native handle (frees device memory)."""
    for ent in _HANDLES.values():
        ent["handle"].close()
    _HANDLES.clear()


class ModelPlanner:
    """Interface of src/mbrl/planners.py:14-25."""

    @staticmethod
    def plan(initial_state, model, cost, sample_action, horizon, initial_trajectory=None, **kwargs):
        raise NotImplementedError


def _host_sample(sample_action, n, horizon, act_dim):
    """One sampler call for all H*N rows, exactly as the reference makes it
    (src/mbrl/planners.py:200); the result is injected into the device rollout."""
    acts = sample_action(batch_size=n * horizon)
    acts = native._f32(acts)
    if acts.shape != (n * horizon, act_dim):
        raise ValueError(f"sample_action returned shape {acts.shape}, expected {(n * horizon, act_dim)}")
    return acts


class RandomShootingPlanner(ModelPlanner):
    """GPU random shooting: N candidate sequences rolled through the dynamics MLP for
    `horizon` steps, scored, argmin (first minimum) returned -- src/mbrl/planners.py:140-216.

    kwargs: num_trajectories (default 1000, planners.py:141); return_states=False skips the fp32
    replay that yields the predicted states (MPCPolicy only uses the first action); sampler="device" (Philox uniform
    on the GPU with the reference sampler's bounds) or "host" (call `sample_action` once on the
    host like the reference and inject the draws: bit-faithful to a given numpy stream);
    engine="fp32"|"bf16"|"fp16"; seed; device."""

    defaults = dict(num_trajectories=1000, sampler="device", engine="fp32", seed=None, device=0, return_states=True)

    @staticmethod
    def plan(initial_state, model, cost, sample_action, horizon, initial_trajectory=None, **kwargs):
        d = RandomShootingPlanner.defaults
        n = int(kwargs.get("num_trajectories", d["num_trajectories"]))
        sampler = kwargs.get("sampler", d["sampler"])
        ent, prob = _get_handle(model, cost, sample_action, horizon, n, 1, kwargs.get("engine", d["engine"]),
                                kwargs.get("device", d["device"]))
        h = ent["handle"]
        seed = kwargs.get("seed", d["seed"])
        if seed is None:
            seed = _default_seed(ent)
        ent["calls"] += 1
        if sampler == "host":
            out = h.plan(initial_state, 1, 1, native.SAMPLE_INJECT_ACTIONS, seed,
                         injected=_host_sample(sample_action, n, horizon, prob.act_dim),
                         actions_only=not kwargs.get("return_states", d["return_states"]))
        elif sampler == "device":
            out = h.plan(initial_state, 1, 1, native.SAMPLE_UNIFORM, seed,
                         actions_only=not kwargs.get("return_states", d["return_states"]))
        else:
            raise ValueError("sampler must be 'device' or 'host'")
        return _as_torch(out["states"][0]), _as_torch(out["actions"][0])


class CEMPlanner(ModelPlanner):
    """Cross-entropy-method planner (not in the reference; SURVEY.md section 8a last row):
    per iteration sample a ~ clip(N(mu[h], sd[h])), roll out + score (same fused kernel as
    random shooting), keep the k = int(elite_frac*N) cheapest (ties -> lower index), refit
    mu/sd per (h, a) with the population std; return the best-ever candidate.

    Warm start across MPC steps (`warm_start`).  MPCPolicy passes the previous plan
    as `initial_trajectory` (None on the first step of an episode, src/mbrl/agents.py:38-47):
      "trajectory" (default)  the mean is seeded with that action sequence as the reference hands it
                              over (un-shifted: agents.py:45 slices the states but not the actions);
      "shift_mean"            the mean is the previous call's FINAL CEM mean shifted by one step
                              (last step repeated) -- the usual MPC-CEM warm start.  The mean stays
                              RESIDENT ON THE DEVICE between calls (MbrlPlanArgs.warm_start) and is
                              dropped when `initial_trajectory` is None (episode start);
      "none"                  always start from the action-range midpoint.
    `init_std` sets the std of a warm-started distribution (default: half the action range)."""

    defaults = dict(num_trajectories=16384, num_iterations=5, elite_frac=0.1, engine="fp16", seed=None, device=0,
                    return_mean=False, init_std=None, return_states=True, warm_start="trajectory")

    @staticmethod
    def plan(initial_state, model, cost, sample_action, horizon, initial_trajectory=None, **kwargs):
        d = CEMPlanner.defaults
        n = int(kwargs.get("num_trajectories", d["num_trajectories"]))
        iters = int(kwargs.get("num_iterations", d["num_iterations"]))
        k = int(kwargs.get("num_elites", max(1, int(kwargs.get("elite_frac", d["elite_frac"]) * n))))
        ent, prob = _get_handle(model, cost, sample_action, horizon, n, iters, kwargs.get("engine", d["engine"]),
                                kwargs.get("device", d["device"]))
        h = ent["handle"]
        seed = kwargs.get("seed", d["seed"])
        if seed is None:
            seed = _default_seed(ent)
        ent["calls"] += 1
        warm = kwargs.get("warm_start", d["warm_start"])
        if warm not in ("trajectory", "shift_mean", "none"):
            raise ValueError("warm_start must be 'trajectory', 'shift_mean' or 'none'")
        mu0 = sd0 = None
        init_std = kwargs.get("init_std", d["init_std"])
        warm_flags = 0
        if warm == "shift_mean":
            # device-resident: the handle keeps the final mean of this plan (KEEP) and seeds the next call
            # with it, shifted by one step on the device (USE) -- no D2H / H2D of the mean between MPC
            # steps.  initial_trajectory is None at the start of an episode (agents.py:38-40): cold start.
            warm_flags = native.WARM_KEEP | (native.WARM_USE if initial_trajectory is not None else 0)
        elif warm == "trajectory" and initial_trajectory is not None:
            acts = native._f32(_stack(initial_trajectory[1])).reshape(-1, prob.act_dim)
            mu0 = np.zeros((horizon, prob.act_dim), np.float32) + 0.5 * (prob.act_lo + prob.act_hi)
            m = min(horizon, acts.shape[0])
            mu0[:m] = acts[:m]
            sd0 = np.full((horizon, prob.act_dim), 0.5 * (prob.act_hi - prob.act_lo) if init_std is None else init_std,
                          np.float32)
        noise = kwargs.get("noise")  # [I, H*N, A] recorded N(0,1) draws (parity runs)
        mode = native.SAMPLE_GAUSSIAN if noise is None else native.SAMPLE_INJECT_NOISE
        out = h.plan(initial_state, iters, k, mode, seed, injected=noise, mu0=mu0, sd0=sd0,
                     return_mean=kwargs.get("return_mean", d["return_mean"]),
                     actions_only=not kwargs.get("return_states", d["return_states"]),
                     warm_start=warm_flags, warm_std=0.0 if init_std is None else float(init_std))
        return _as_torch(out["states"][0]), _as_torch(out["actions"][0])


class GradientDescentPlanner(ModelPlanner):
    """GPU GradientDescentPlanner (src/mbrl/planners.py:28-137): Adam(lr 0.01) on the action sequence with
    back-propagation through the H-step model rollout, `num_iterations` (40) iterations at most, stopping when
    mean|a_old - a_new| < `stop_condition` (0.002).  Returns, like the reference, H+1 states (s_0 first; the
    states of the last forward pass) and H actions as lists of [1, .] tensors.

    Batched: `num_restarts` > 1 optimises that many independently drawn initial sequences at once (one CTA
    each) and returns the one whose last forward pass had the lowest loss; restart 0 is always the
    reference's own start (initial_trajectory's actions, or one sample_action(batch_size=horizon) call)."""

    defaults = dict(num_iterations=40, stop_condition=0.002, num_restarts=1, lr=0.01, engine="fp32", device=0)

    @staticmethod
    def plan(initial_state, model, cost, sample_action, horizon, initial_trajectory=None, **kwargs):
        import torch
        d = GradientDescentPlanner.defaults
        ent, prob = _get_handle(model, cost, sample_action, horizon, 1, 1, kwargs.get("engine", d["engine"]),
                                kwargs.get("device", d["device"]))
        restarts = int(kwargs.get("num_restarts", d["num_restarts"]))
        if initial_trajectory is not None:
            first = native._f32(_stack(initial_trajectory[1])).reshape(-1, prob.act_dim)[:horizon]
        else:
            first = native._f32(sample_action(batch_size=horizon)).reshape(horizon, prob.act_dim)
        init = [first] + [native._f32(sample_action(batch_size=horizon)).reshape(horizon, prob.act_dim) for _ in range(restarts - 1)]
        out = ent["handle"].plan_gd(initial_state, np.stack(init), kwargs.get("num_iterations", d["num_iterations"]),
                                    kwargs.get("stop_condition", d["stop_condition"]), kwargs.get("lr", d["lr"]))
        ent["calls"] += 1
        best = int(np.argmin(out["cost"])) if restarts > 1 else 0
        states, actions = torch.from_numpy(out["states"][best]), torch.from_numpy(out["actions"][best])
        return list(states.split(1, 0)), list(actions.split(1, 0))


def _stack(seq):
    import torch
    if isinstance(seq, (list, tuple)):
        return torch.stack([torch.as_tensor(s).reshape(-1) for s in seq])
    return seq
