"""Turns the reference's opaque planner callables into an explicit problem description.

``ModelPlanner.plan`` receives ``model``, ``cost`` and ``sample_action`` as callables
(src/mbrl/planners.py:16-25).  In the reference they are ``functools.partial`` objects built
by ``GoalStateAgent.__init__`` (src/mbrl/agents.py:219-235):

    model         = partial(<Model nn.Module>, normalize_state=partial(normalize_field, field_name="observations",
                            stats=D), normalize_action=partial(..."actions"...), unnormalize_state=partial(...))
    cost          = partial(state_action_cost, state_cost=<SmoothAbsLoss>, action_cost=<CoshLoss>)
    sample_action = partial(EnvWrapper._sample_action, action_spec=spec)

A GPU planner cannot call Python callables per candidate, so this module introspects those
partials (or accepts an explicit ``PlanningProblem``).  Anything it does not recognise raises
``TypeError`` -- there is deliberately no fallback to calling the callables on the CPU.
"""
from __future__ import annotations

import functools
from dataclasses import dataclass
from typing import Any, Optional, Tuple

import numpy as np


@dataclass
class PlanningProblem:
    """Explicit description of one planning problem (host arrays / tensors)."""

    W1: Any
    b1: Any
    W2: Any
    b2: Any
    W3: Any
    b3: Any
    mu_s: Any = None
    sd_s: Any = None
    mu_a: Any = None
    sd_a: Any = None
    cost_w: Any = None
    goal: Any = None
    alpha: float = 0.4
    beta: float = 0.25
    act_lo: float = -1.0
    act_hi: float = 1.0
    cost_kind: int = 0  # native.COST_*: 0 = SmoothAbs + Cosh (the reference's), 1 = dm_control cartpole swing-up,
    #                     2 = ModelWithReward's reward head (RewardAgent)
    W4: Any = None      # [1, U] linear4.weight / [1] bias / "rewards" statistics -- cost_kind 2 only
    b4: Any = None
    mu_r: Any = 0.0
    sd_r: Any = 1.0

    @property
    def obs_dim(self) -> int:
        return int(self.W3.shape[0])

    @property
    def act_dim(self) -> int:
        return int(self.W1.shape[1]) - self.obs_dim

    @property
    def hidden(self) -> int:
        return int(self.W1.shape[0])


def _stats_of(norm_partial, what: str) -> Tuple[Any, Any, Any]:
    """(mean, std, identity-token) from a partial(normalize_field|unnormalize_field, field_name=, stats=)."""
    if norm_partial is None:
        return None, None, None
    if not isinstance(norm_partial, functools.partial):
        raise TypeError(f"{what}: expected functools.partial(normalize_field, field_name=..., stats=...), got {type(norm_partial)!r}")
    kw = norm_partial.keywords or {}
    if "stats" not in kw or "field_name" not in kw:
        raise TypeError(f"{what}: partial lacks 'stats'/'field_name' keywords (src/mbrl/agents.py:215-217)")
    entry = kw["stats"][kw["field_name"]]
    return entry["mean"], entry["std"], entry


_LINEAR_EMBEDDINGS: dict = {}


def embed_linear_model(weight, bias):
    """LinearModel (src/mbrl/models.py:113-122) is a single Linear(D, O).  It runs on the same
    kernels as the 3-layer MLP through an exact ReLU embedding with hidden width 2D:
        h1 = relu([x; -x])          W1 = [I; -I], b1 = 0
        h2 = relu(I h1) = h1        W2 = I,       b2 = 0      (h1 >= 0 already)
        y  = [W, -W] h2 + b         = W relu(x) - W relu(-x) + b = W x + b
    Every product W[i,j]*x[j] appears once with its exact value (the other half contributes exact
    zeros), so only the summation order differs from the reference's dot product.
    Returns (W1, b1, W2, b2, W3, b3) as float32 numpy arrays; cached per (tensor, version)."""
    key = (id(weight), int(getattr(weight, "_version", 0)), id(bias), int(getattr(bias, "_version", 0)))
    hit = _LINEAR_EMBEDDINGS.get(key)
    if hit is not None:
        return hit
    W = np.ascontiguousarray(weight.detach().cpu().numpy() if hasattr(weight, "detach") else weight, np.float32)
    b = np.ascontiguousarray(bias.detach().cpu().numpy() if hasattr(bias, "detach") else bias, np.float32)
    D = W.shape[1]
    eye = np.eye(D, dtype=np.float32)
    out = (np.concatenate([eye, -eye], 0), np.zeros(2 * D, np.float32), np.eye(2 * D, dtype=np.float32),
           np.zeros(2 * D, np.float32), np.concatenate([W, -W], 1), b)
    _LINEAR_EMBEDDINGS.clear()  # one live model is the normal case; do not accumulate stale versions
    _LINEAR_EMBEDDINGS[key] = out
    return out


class _Layer:
    """Minimal stand-in for nn.Linear carrying embedded weights."""

    def __init__(self, weight, bias):
        self.weight, self.bias = weight, bias


def _linear_layers(module):
    names = ("linear1", "linear2", "linear3")
    if hasattr(module, "linear1") and not hasattr(module, "linear2") and not hasattr(module, "linear3"):
        if getattr(module, "noise", None) is not None:
            raise TypeError("LinearModel(noise=...) adds fresh Gaussian noise per forward (src/mbrl/models.py:122); unsupported")
        W1, b1, W2, b2, W3, b3 = embed_linear_model(module.linear1.weight, module.linear1.bias)
        return [_Layer(W1, b1), _Layer(W2, b2), _Layer(W3, b3)]
    if not all(hasattr(module, n) for n in names):
        raise TypeError(
            f"{type(module).__name__}: unsupported dynamics model (need linear1/linear2/linear3 as in "
            "src/mbrl/models.py:96-110); pass an explicit PlanningProblem instead")
    if getattr(module, "noise", None) is not None:
        raise TypeError("Model(noise=...) adds fresh Gaussian noise per forward (src/mbrl/models.py:110); unsupported")
    return [getattr(module, n) for n in names]


def _composed(fn):
    """RewardAgent wraps the wired ModelWithReward as compose(partial(...), itemgetter(i))
    (src/mbrl/agents.py:290-295, 349-358): a closure over (a, b).  Returns (partial, index) or None."""
    import operator
    cells = getattr(fn, "__closure__", None)
    if not cells:
        return None
    inner = index = None
    for c in cells:
        v = c.cell_contents
        if isinstance(v, functools.partial):
            inner = v
        elif isinstance(v, operator.itemgetter):
            probe = v(("state", "reward"))
            index = 0 if probe == "state" else 1 if probe == "reward" else None
    return (inner, index) if inner is not None and index is not None else None


def _action_bounds(sample_action):
    lo, hi = -1.0, 1.0
    if isinstance(sample_action, functools.partial) and "action_spec" in (sample_action.keywords or {}):
        spec = sample_action.keywords["action_spec"]
        # EnvWrapper._sample_action: dimension 0's bounds for every dim, clipped to +-3
        # (src/mbrl/env_wrappers.py:52-55)
        lo, hi = max(float(spec.minimum[0]), -3.0), min(float(spec.maximum[0]), 3.0)
    return lo, hi


def _reward_problem(model, cost, sample_action):
    """The RewardAgent wiring: model = compose(wired, itemgetter(0)), cost = compose(wired, itemgetter(1))."""
    from . import native
    (wired_m, i_m), (wired_c, i_c) = _composed(model), _composed(cost)
    if (i_m, i_c) != (0, 1) or wired_m.func is not wired_c.func or not hasattr(wired_m.func, "linear4"):
        raise TypeError("composed model/cost must be RewardAgent's (ModelWithReward, itemgetter(0)/(1)) pair "
                        "(src/mbrl/agents.py:349-358)")
    module = wired_m.func
    l1, l2, l3 = _linear_layers(module)
    l4 = module.linear4
    kw = wired_m.keywords or {}
    mu_s, sd_s, tok_s = _stats_of(kw.get("normalize_state"), "normalize_state")
    mu_u, sd_u, tok_u = _stats_of(kw.get("unnormalize_state"), "unnormalize_state")
    mu_a, sd_a, _ = _stats_of(kw.get("normalize_action"), "normalize_action")
    mu_r, sd_r, _ = _stats_of(kw.get("unnormalize_reward"), "unnormalize_reward")
    if (tok_s is None) != (tok_u is None) or (tok_s is not None and tok_s is not tok_u):
        raise TypeError("normalize_state and unnormalize_state must use the same statistics entry")
    lo, hi = _action_bounds(sample_action)
    prob = PlanningProblem(
        W1=l1.weight, b1=l1.bias, W2=l2.weight, b2=l2.bias, W3=l3.weight, b3=l3.bias,
        mu_s=mu_s, sd_s=sd_s, mu_a=mu_a, sd_a=sd_a, act_lo=lo, act_hi=hi,
        cost_kind=native.COST_REWARD_HEAD, W4=l4.weight, b4=l4.bias,
        mu_r=0.0 if mu_r is None else mu_r, sd_r=1.0 if sd_r is None else sd_r,
    )
    params = [l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias, l4.weight, l4.bias]
    fp = (
        tuple((id(t), int(getattr(t, "_version", 0))) for t in params),
        tuple((id(t), int(getattr(t, "_version", 0))) for t in (mu_s, sd_s, mu_a, sd_a, mu_r, sd_r) if t is not None),
        "reward_head", lo, hi,
    )
    return prob, fp


def problem_from_callables(model, cost, sample_action) -> Tuple[PlanningProblem, tuple]:
    """Introspect the reference's partials.  Returns (problem, fingerprint); the fingerprint
    changes whenever the host retrains the model in place (param._version), replaces the
    statistics entries (src/mbrl/data.py:244-249) or moves the goal (models.py:240-241)."""
    if isinstance(model, PlanningProblem):
        return model, ("explicit", id(model))
    if _composed(model) is not None and _composed(cost) is not None:
        return _reward_problem(model, cost, sample_action)
    if isinstance(model, functools.partial) and hasattr(model.func, "linear4"):
        raise TypeError("ModelWithReward must be wired as RewardAgent does (compose(..., itemgetter(0)) as model and "
                        "compose(..., itemgetter(1)) as cost, src/mbrl/agents.py:349-358)")
    if not isinstance(model, functools.partial) or not hasattr(model.func, "parameters"):
        raise TypeError(
            "model must be functools.partial(<nn.Module>, normalize_state=..., normalize_action=..., "
            "unnormalize_state=...) as built in src/mbrl/agents.py:225-230, or a PlanningProblem; "
            "opaque callables cannot run on the GPU and there is no CPU fallback")
    module = model.func
    l1, l2, l3 = _linear_layers(module)
    kw = model.keywords or {}
    mu_s, sd_s, tok_s = _stats_of(kw.get("normalize_state"), "normalize_state")
    mu_u, sd_u, tok_u = _stats_of(kw.get("unnormalize_state"), "unnormalize_state")
    mu_a, sd_a, tok_a = _stats_of(kw.get("normalize_action"), "normalize_action")
    if (tok_s is None) != (tok_u is None) or (tok_s is not None and tok_s is not tok_u):
        raise TypeError("normalize_state and unnormalize_state must use the same statistics entry")

    if not isinstance(cost, functools.partial):
        raise TypeError("cost must be functools.partial(state_action_cost, state_cost=SmoothAbsLoss, action_cost=CoshLoss) "
                        "(src/mbrl/agents.py:231)")
    ckw = cost.keywords or {}
    sc, ac = ckw.get("state_cost"), ckw.get("action_cost")
    if sc is None or ac is None or not all(hasattr(sc, a) for a in ("weights", "goal_state", "alpha")) or not hasattr(ac, "alpha"):
        raise TypeError("cost partial must carry state_cost (weights, goal_state, alpha) and action_cost (alpha)")

    lo, hi = _action_bounds(sample_action)

    prob = PlanningProblem(
        W1=l1.weight, b1=l1.bias, W2=l2.weight, b2=l2.bias, W3=l3.weight, b3=l3.bias,
        mu_s=mu_s, sd_s=sd_s, mu_a=mu_a, sd_a=sd_a,
        cost_w=sc.weights, goal=sc.goal_state, alpha=float(sc.alpha), beta=float(ac.alpha),
        act_lo=lo, act_hi=hi,
    )
    params = list(module.parameters())  # the live tensors (the embedded copies of a LinearModel are derived from them)
    fp = (
        tuple((id(t), int(getattr(t, "_version", 0))) for t in params),
        tuple((id(t), int(getattr(t, "_version", 0))) for t in (mu_s, sd_s, mu_a, sd_a, sc.weights, sc.goal_state) if t is not None),
        float(sc.alpha), float(ac.alpha), lo, hi,
    )
    return prob, fp


# ---- per-call fast path ------------------------------------------------------------------------
# MPCPolicy hands the SAME partial objects to every plan() call (src/mbrl/agents.py:48-55), so the
# introspection above (40 us of Python) is done once per (model, cost, sample_action) triple; later
# calls only re-check what can change underneath those objects: in-place training bumps
# param._version (models.py:84-86), add_rollouts REPLACES the statistics entries (data.py:244-249),
# set_goal_state replaces goal_state (models.py:240-241).
_FAST: dict = {}


def _probe_of(model, cost):
    """A cheap callable whose value changes whenever problem_from_callables' fingerprint would."""
    module = model.func
    params = list(module.parameters())
    norm = [p for p in ((model.keywords or {}).get(k) for k in ("normalize_state", "normalize_action", "unnormalize_state"))
            if p is not None]
    stat_refs = [(p.keywords["stats"], p.keywords["field_name"]) for p in norm]
    sc, ac = cost.keywords["state_cost"], cost.keywords["action_cost"]

    def probe():
        out = [t._version for t in params]
        for stats, name in stat_refs:
            entry = stats[name]
            m, s = entry["mean"], entry["std"]
            out += (id(m), getattr(m, "_version", 0), id(s), getattr(s, "_version", 0))
        w, g = sc.weights, sc.goal_state
        out += (id(w), getattr(w, "_version", 0), id(g), getattr(g, "_version", 0), sc.alpha, ac.alpha)
        return out
    return probe


def problem_from_callables_cached(model, cost, sample_action):
    """problem_from_callables with the per-call fast path for the GoalStateAgent wiring; anything else
    (explicit PlanningProblem, RewardAgent closures) takes the full introspection every call."""
    key = (id(model), id(cost), id(sample_action))
    ent = _FAST.get(key)
    if ent is not None and ent[0] is model and ent[1] is cost and ent[2] is sample_action:
        if ent[3]() == ent[4]:
            return ent[5], ent[6]
    prob, fp = problem_from_callables(model, cost, sample_action)
    if isinstance(model, functools.partial) and isinstance(cost, functools.partial) and hasattr(model.func, "parameters") \
            and "state_cost" in (cost.keywords or {}):
        probe = _probe_of(model, cost)
        if len(_FAST) > 64:
            _FAST.clear()
        _FAST[key] = (model, cost, sample_action, probe, probe(), prob, fp)  # strong refs keep the ids from being recycled
    return prob, fp
