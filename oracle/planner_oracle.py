"""torch-CPU fp32 restatement of the reference planning path (TEST INFRASTRUCTURE).

Each function cites the reference lines it follows (paths relative to the reference
root).  The reference computes in fp32 on the CPU with torch ops; this file uses the
same torch ops in the same order so that, on one machine, results are bit-identical to
the reference (checked by tests/test_oracle_golden.py against fixtures generated from
the reference itself).

Layout convention of the reference (src/mbrl/planners.py:199-209): flat *step-major*
buffers, row ``h*N + n`` is candidate ``n`` at step ``h``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# Parameters of one planning problem
# --------------------------------------------------------------------------------------
@dataclass
class PlannerParams:
    """Everything ``DynamicsModel.forward`` + ``state_action_cost`` close over."""

    W1: torch.Tensor  # [U, O+A]  nn.Linear layout (out, in)   src/mbrl/models.py:99
    b1: torch.Tensor  # [U]
    W2: torch.Tensor  # [U, U]                                  src/mbrl/models.py:100
    b2: torch.Tensor  # [U]
    W3: torch.Tensor  # [O, U]                                  src/mbrl/models.py:101
    b3: torch.Tensor  # [O]
    mu_s: torch.Tensor  # [O]  stats["observations"]["mean"]    src/mbrl/data.py:262-269
    sd_s: torch.Tensor  # [O]
    mu_a: torch.Tensor  # [A]
    sd_a: torch.Tensor  # [A]
    cost_w: torch.Tensor  # [O]  SmoothAbsLoss.weights          src/mbrl/models.py:249-253
    goal: torch.Tensor  # [O]    SmoothAbsLoss.goal_state
    alpha: float = 0.4  #        SmoothAbsLoss.alpha            src/mbrl/models.py:249
    beta: float = 0.25  #        CoshLoss.alpha                 src/mbrl/models.py:267
    act_lo: float = -1.0
    act_hi: float = 1.0
    # ModelWithReward's reward head and the "rewards" statistics (src/mbrl/models.py:125-163,
    # src/mbrl/agents.py:342-366); None for the plain dynamics Model
    W4: Optional[torch.Tensor] = None  # [1, U]   linear4.weight
    b4: Optional[torch.Tensor] = None  # [1]
    mu_r: float = 0.0                   # stats["rewards"]["mean"]
    sd_r: float = 1.0

    # LinearModel (src/mbrl/models.py:113-122): W2/b2/W3/b3 are None and W1 is the single Linear(D, O)

    @property
    def obs_dim(self) -> int:
        return int((self.W3 if self.W3 is not None else self.W1).shape[0])

    @property
    def act_dim(self) -> int:
        return int(self.W1.shape[1]) - self.obs_dim

    @property
    def hidden(self) -> int:
        return int(self.W1.shape[0])


def synthetic_params(obs_dim: int, act_dim: int, hidden: int, seed: int = 0) -> PlannerParams:
    """Synthetic problem of SURVEY.md section 8(d): default ``nn.Linear`` init
    (Kaiming-uniform a=sqrt(5), the initialiser ``Model.__init__`` gets,
    src/mbrl/models.py:99-101), mu_s~N(0,1), sd_s~U(0.5,1.5), mu_a=0, sd_a=1/sqrt(3),
    SmoothAbs w=1 g=0 alpha=0.4, Cosh beta=0.25, action bounds +-1."""
    g = torch.Generator().manual_seed(seed)

    def linear(out_f: int, in_f: int) -> Tuple[torch.Tensor, torch.Tensor]:
        bound = 1.0 / float(np.sqrt(in_f))
        W = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
        b = (torch.rand(out_f, generator=g) * 2 - 1) * bound
        return W, b

    W1, b1 = linear(hidden, obs_dim + act_dim)
    W2, b2 = linear(hidden, hidden)
    W3, b3 = linear(obs_dim, hidden)
    mu_s = torch.randn(obs_dim, generator=g)
    sd_s = torch.rand(obs_dim, generator=g) + 0.5
    return PlannerParams(
        W1, b1, W2, b2, W3, b3,
        mu_s, sd_s,
        torch.zeros(act_dim), torch.full((act_dim,), float(1.0 / np.sqrt(3.0))),
        torch.ones(obs_dim), torch.zeros(obs_dim),
    )


def synthetic_state(p: PlannerParams, call: int = 0) -> torch.Tensor:
    """s0 ~ N(mu_s, sd_s), seed 1000+call."""
    g = torch.Generator().manual_seed(1000 + call)
    return p.mu_s + p.sd_s * torch.randn(p.obs_dim, generator=g)


# --------------------------------------------------------------------------------------
# Model / normalisers / cost
# --------------------------------------------------------------------------------------
def normalize(x: torch.Tensor, mean: torch.Tensor, std: torch.Tensor) -> torch.Tensor:
    """TransitionsDataset.normalize_field, src/mbrl/data.py:258-260 (std unguarded)."""
    return (x - mean) / std


def unnormalize(x: torch.Tensor, mean: torch.Tensor, std: torch.Tensor) -> torch.Tensor:
    """TransitionsDataset.unnormalize_field, src/mbrl/data.py:255-257."""
    return (x * std) + mean


def mlp_forward(p: PlannerParams, x: torch.Tensor) -> torch.Tensor:
    """Model._forward, src/mbrl/models.py:106-110 (noise=None): relu(L1), relu(L2), L3."""
    if p.W2 is None:  # LinearModel._forward, src/mbrl/models.py:119-122
        return F.linear(x, p.W1, p.b1)
    h = torch.relu(F.linear(x, p.W1, p.b1))
    h = torch.relu(F.linear(h, p.W2, p.b2))
    return F.linear(h, p.W3, p.b3)


def dynamics_forward(p: PlannerParams, state: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
    """DynamicsModel.forward, src/mbrl/models.py:13-29: normalise action and state,
    concatenate [state, action], MLP, un-normalise.  Predicts the next state directly
    (no residual add)."""
    a = normalize(action, p.mu_a, p.sd_a)
    s = normalize(state, p.mu_s, p.sd_s)
    y = mlp_forward(p, torch.cat([s, a], dim=1))
    return unnormalize(y, p.mu_s, p.sd_s)


def smooth_abs_cost(p: PlannerParams, state: torch.Tensor) -> torch.Tensor:
    """SmoothAbsLoss.forward, src/mbrl/models.py:255-259: sum over obs dims."""
    x = state - p.goal
    return torch.sum(torch.sqrt((x * p.cost_w) ** 2 + p.alpha ** 2) - p.alpha, dim=-1)


def cosh_cost(p: PlannerParams, action: torch.Tensor) -> torch.Tensor:
    """CoshLoss.forward, src/mbrl/models.py:271-272: mean over action dims."""
    return (p.beta ** 2) * torch.mean(torch.cosh(action / p.beta) - 1, dim=-1)


def state_action_cost(p: PlannerParams, state: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
    """state_action_cost, src/mbrl/agents.py:182-183."""
    return smooth_abs_cost(p, state) + cosh_cost(p, action)


# --------------------------------------------------------------------------------------
# Random shooting
# --------------------------------------------------------------------------------------
def reward_head_cost(p: PlannerParams, state: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
    """RewardAgent's cost callable: compose(partial(ModelWithReward, normalisers...), itemgetter(1))
    (src/mbrl/agents.py:349-358): a SECOND trunk evaluation at the state it is handed (the planner
    hands it the predicted next states, planners.py:210) with the same action, through the reward
    head, un-normalised with the reward statistics (models.py:135-163).  The planner minimises it
    as it would a cost (the reference's behaviour, kept).  Returns [B, 1] like the reference."""
    x = torch.cat([normalize(state, p.mu_s, p.sd_s), normalize(action, p.mu_a, p.sd_a)], dim=1)
    h1 = torch.relu(torch.nn.functional.linear(x, p.W1, p.b1))
    h2 = torch.relu(torch.nn.functional.linear(h1, p.W2, p.b2))
    r = torch.nn.functional.linear(h2, p.W4, p.b4)
    return r * p.sd_r + p.mu_r  # unnormalize_field, data.py:255-257


def rollout_costs(
    p: PlannerParams, s0: torch.Tensor, actions: torch.Tensor, horizon: int, n: int, cost: str = "goal"
) -> Tuple[torch.Tensor, np.ndarray]:
    """The hot loop of RandomShootingPlanner._generate_trajectories,
    src/mbrl/planners.py:199-210, for the dynamics MLP + SmoothAbs/Cosh cost.

    actions: [H*N, A] step-major.  Returns (states [H*N, O] step-major -- the predicted
    s_1..s_H, s_0 not included -- and costs [N] float32).  Row cost of step h pairs
    s_{h+1} with a_h; trajectory cost is the sum over h."""
    with torch.no_grad():
        states = torch.zeros((n * horizon, s0.shape[0]))
        for h in range(horizon):
            if h == 0:
                cur = s0.unsqueeze(0).repeat_interleave(n, dim=0)
            else:
                cur = states[(h - 1) * n: h * n]
            states[h * n: (h + 1) * n] = dynamics_forward(p, cur, actions[h * n: (h + 1) * n])
        if cost == "reward_head":
            costs = reward_head_cost(p, states, actions).view(horizon, n).sum(0).numpy()
        else:
            costs = state_action_cost(p, states, actions).view(horizon, n).sum(0).numpy()
    return states, costs


def rs_plan(
    p: PlannerParams, s0: torch.Tensor, actions: torch.Tensor, horizon: int, n: int
) -> Dict[str, object]:
    """RandomShootingPlanner._plan, src/mbrl/planners.py:166-187: np.argmin (first
    minimum on ties) over the trajectory costs; returns that candidate's [H,O] states
    and [H,A] actions."""
    states, costs = rollout_costs(p, s0, actions, horizon, n)
    idx = int(np.argmin(costs))
    return dict(
        idx=idx,
        costs=costs,
        states=states.view(horizon, n, -1)[:, idx].clone(),
        actions=actions.view(horizon, n, -1)[:, idx].clone(),
    )


def reference_style_generate(
    initial_state: torch.Tensor,
    model: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
    cost: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
    sample_action: Callable[..., torch.Tensor],
    horizon: int,
    num_trajectories: int,
):
    """Callable-level restatement of RandomShootingPlanner._generate_trajectories
    (src/mbrl/planners.py:189-216) *including* its bookkeeping: autograd left on, one
    sampler call for all H*N actions, and the per-candidate Python list of views
    (planners.py:211-215).  Used for the ring-world known answer and as the faithful
    CPU baseline timed by bench.py."""
    n, hz = num_trajectories, horizon
    s_all = torch.zeros((n * hz, initial_state.shape[0]))
    a_all = sample_action(batch_size=n * hz)
    for h in range(hz):
        prev = (
            initial_state.unsqueeze(dim=0).repeat_interleave(n, dim=0)
            if h == 0
            else s_all[(h - 1) * n: h * n]
        )
        s_all[h * n: (h + 1) * n] = model(prev, a_all[h * n: (h + 1) * n])
    per_traj = cost(s_all, a_all).view(hz, n).sum(0).detach().numpy()
    s_view = s_all.view((hz, n, -1))
    a_view = a_all.view((hz, n, -1))
    trajectories = [(s_view[:, i], a_view[:, i]) for i in range(n)]
    return trajectories, per_traj


def reference_style_plan(initial_state, model, cost, sample_action, horizon, num_trajectories=1000):
    """RandomShootingPlanner.plan/_plan, src/mbrl/planners.py:143-187."""
    trajs, costs = reference_style_generate(
        initial_state, model, cost, sample_action, horizon, num_trajectories
    )
    return trajs[int(np.argmin(costs))]


def params_as_callables(p: PlannerParams):
    """(model, cost) callables with the reference's call signature, for
    ``reference_style_plan``."""
    return (lambda s, a: dynamics_forward(p, s, a)), (lambda s, a: state_action_cost(p, s, a))


# --------------------------------------------------------------------------------------
# Elite selection / refit / CEM  (no reference lines -- see SURVEY.md section 8c)
# --------------------------------------------------------------------------------------
def topk_stable(costs: np.ndarray, k: int) -> np.ndarray:
    """k smallest costs, ties broken toward the lower index, ascending (cost, index)
    order; consistent at k=1 with np.argmin (src/mbrl/planners.py:184).  NaN costs sort
    last (numpy's convention)."""
    return np.argsort(np.asarray(costs), kind="stable")[:k].astype(np.int64)


def refit(actions_hna: torch.Tensor, elite: np.ndarray) -> Tuple[torch.Tensor, torch.Tensor]:
    """mean/std (population std, unbiased=False) of the elite action sequences per
    (h, a) column.  actions_hna: [H, N, A]."""
    e = actions_hna[:, torch.as_tensor(elite, dtype=torch.long)]
    return e.mean(dim=1), e.std(dim=1, unbiased=False)


def gaussian_actions(
    mu: torch.Tensor, sd: torch.Tensor, z: torch.Tensor, n: int, lo: float, hi: float
) -> torch.Tensor:
    """clip(mu[h] + sd[h] * z[h*N+n], lo, hi) in the step-major layout.
    mu, sd: [H, A]; z: [H*N, A]."""
    m = mu.repeat_interleave(n, dim=0)
    s = sd.repeat_interleave(n, dim=0)
    return torch.clamp(m + s * z, min=lo, max=hi)


def cem_plan(
    p: PlannerParams,
    s0: torch.Tensor,
    noise: torch.Tensor,  # [I, H*N, A] standard normal draws (recorded / injected)
    horizon: int,
    n: int,
    k: int,
    mu0: Optional[torch.Tensor] = None,
    sd0: Optional[torch.Tensor] = None,
) -> Dict[str, object]:
    """Reference-composed CEM: every iteration's rollout + cost
    is ``rollout_costs`` (= the reference's _generate_trajectories with a Gaussian
    sampler); elite select, refit and best-ever tracking are this repo's definition.

    Returns the best-ever candidate (earliest iteration wins ties, then lowest index)."""
    A = p.act_dim
    mu = torch.full((horizon, A), 0.5 * (p.act_lo + p.act_hi)) if mu0 is None else mu0.clone()
    sd = torch.full((horizon, A), 0.5 * (p.act_hi - p.act_lo)) if sd0 is None else sd0.clone()
    best = dict(cost=np.float32(np.inf), it=-1, idx=-1, states=None, actions=None)
    hist: List[Dict[str, object]] = []
    for it in range(noise.shape[0]):
        acts = gaussian_actions(mu, sd, noise[it], n, p.act_lo, p.act_hi)
        states, costs = rollout_costs(p, s0, acts, horizon, n)
        elite = topk_stable(costs, k)
        j = int(elite[0])
        if costs[j] < best["cost"]:
            best = dict(
                cost=costs[j], it=it, idx=j,
                states=states.view(horizon, n, -1)[:, j].clone(),
                actions=acts.view(horizon, n, -1)[:, j].clone(),
            )
        mu, sd = refit(acts.view(horizon, n, -1), elite)
        hist.append(dict(costs=costs, elite=elite, mu=mu.clone(), sd=sd.clone(), actions=acts))
    return dict(best=best, mu=mu, sd=sd, history=hist)


# --------------------------------------------------------------------------------------
# GradientDescentPlanner (src/mbrl/planners.py:28-137)
# --------------------------------------------------------------------------------------
def gd_plan(p: PlannerParams, s0: torch.Tensor, init_actions: torch.Tensor, horizon: int, num_iterations: int = 40,
            stop_condition: float = 0.002, lr: float = 0.01):
    """Restates GradientDescentPlanner._optimize_trajectory (planners.py:101-137) on the explicit problem:
    Adam(lr) on the [H, A] action tensor, loss = sum over steps of state_action_cost(s_{h+1}, a_h) with
    s_{h+1} = DynamicsModel.forward(s_h, a_h) (full back-propagation through time), stop as soon as
    mean|a_old - a_new| < stop_condition.  Like the reference it returns the states of the LAST forward
    pass (computed with the actions before the final Adam step, planners.py:121-135) together with the
    updated actions.  Returns (states [H+1, O], actions [H, A], iterations run)."""
    states = torch.zeros((horizon + 1, s0.shape[-1]))
    states[0] = s0
    actions = init_actions.clone().detach().reshape(horizon, -1)
    actions.requires_grad = True
    opt = torch.optim.Adam([actions], lr=lr)
    ran = 0
    for _ in range(num_iterations):
        opt.zero_grad()
        for i in range(horizon):
            states[i + 1] = dynamics_forward(p, states[i:i + 1], actions[i:i + 1])
        loss = torch.sum(state_action_cost(p, states[1:], actions))
        loss.backward(retain_graph=True)
        old = actions.clone().detach()
        opt.step()
        ran += 1
        change = torch.mean(torch.abs(old - actions)).detach().numpy()
        if change < stop_condition:
            break
    return states.detach().clone(), actions.detach().clone(), ran
