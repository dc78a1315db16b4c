"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF.

Only works where the read-only reference checkout exists (``/root/reference`` in the build
container; it does not exist on the GPU box, which is why the outputs are committed).

    python tests/golden/make_golden.py

What is imported from the reference (unmodified, from where it lies):
  src/mbrl/planners.py   RandomShootingPlanner.plan / _generate_trajectories
  src/mbrl/models.py     Model, LinearModel, ModelWithReward, SmoothAbsLoss, CoshLoss
  src/mbrl/data.py       TransitionsDataset.normalize_field / unnormalize_field
  src/mbrl/agents.py     MPCPolicy, state_action_cost, compose   (needs the stubs below)
  src/mbrl/env_wrappers.py  EnvWrapper._sample_action        (needs the stubs below)
  dm_control/dm_control/utils/rewards.py  tolerance           (numpy only)

Fixtures: ring_world, rs_cartpole, rs_cheetah_small (random shooting: costs, argmin, plan, first
action through MPCPolicy), cem_cheetah_small, cem_cartpole_small (reference-composed CEM),
rs_reward_head (RewardAgent wiring), rs_linear_model (--model lin), tolerance (rewards.tolerance
grid), humanoid_reward (Humanoid.get_reward composed with the reference's tolerance), locomotion_reward
(Cheetah.get_reward and walker-walk PlanarWalker.get_reward, same construction), gd_small / gd_early_stop /
gd_cheetah (the reference's GradientDescentPlanner run as shipped).

Third-party modules the reference imports at module scope but which are not installed
here (tensorboardX, colorlog, dm_env, dm_control.suite, PIL) are stubbed in sys.modules;
``torch.autograd.gradcheck.zero_gradients`` (removed from torch, imported by
src/mbrl/utils.py:6) is shimmed.  None of them is on the planning path.
"""
import os
import sys
import types
from functools import partial

import numpy as np
import torch

REF = os.environ.get("MBRL_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("tensorboardX", SummaryWriter=type("SummaryWriter", (), {}))
    mod("colorlog", ColoredFormatter=type("ColoredFormatter", (), {"__init__": lambda s, *a, **k: None}))
    mod("dm_env", Environment=type("Environment", (), {}))
    dmc = mod("dm_control")
    dmc.suite = mod("dm_control.suite")
    try:
        import PIL  # noqa: F401
    except Exception:
        pil = mod("PIL")
        pil.Image = mod("PIL.Image")
    import importlib

    gc = importlib.import_module("torch.autograd.gradcheck")  # the module, not the function
    if not hasattr(gc, "zero_gradients"):
        gc.zero_gradients = lambda *a, **k: None
    sys.path.insert(0, REF)


def _reference_problem(obs, act, hidden, seed=0):
    """Build model / cost / stats the way GoalStateAgent.__init__ wires them
    (src/mbrl/agents.py:219-235), with the synthetic statistics of SURVEY.md 8(d)."""
    from src.mbrl.models import Model, SmoothAbsLoss, CoshLoss
    from src.mbrl.data import TransitionsDataset
    from src.mbrl.agents import state_action_cost

    torch.manual_seed(seed)
    net = Model(obs, act, hidden_units=hidden)
    g = torch.Generator().manual_seed(seed + 17)
    stats = {
        "observations": {"mean": torch.randn(obs, generator=g), "std": torch.rand(obs, generator=g) + 0.5},
        "actions": {"mean": 0.1 * torch.randn(act, generator=g), "std": torch.rand(act, generator=g) * 0.3 + 0.4},
    }
    weights = torch.rand(obs, generator=g) + 0.5
    goal = 0.3 * torch.randn(obs, generator=g)
    state_cost = SmoothAbsLoss(weights=weights, goal_state=goal)
    action_cost = CoshLoss()
    model = partial(
        net,
        normalize_state=partial(TransitionsDataset.normalize_field, field_name="observations", stats=stats),
        normalize_action=partial(TransitionsDataset.normalize_field, field_name="actions", stats=stats),
        unnormalize_state=partial(TransitionsDataset.unnormalize_field, field_name="observations", stats=stats),
    )
    cost = partial(state_action_cost, state_cost=state_cost, action_cost=action_cost)
    s0 = stats["observations"]["mean"] + stats["observations"]["std"] * torch.randn(obs, generator=g)
    arrays = dict(
        W1=net.linear1.weight, b1=net.linear1.bias, W2=net.linear2.weight, b2=net.linear2.bias,
        W3=net.linear3.weight, b3=net.linear3.bias,
        mu_s=stats["observations"]["mean"], sd_s=stats["observations"]["std"],
        mu_a=stats["actions"]["mean"], sd_a=stats["actions"]["std"],
        cost_w=weights, goal=goal, s0=s0,
        alpha=torch.tensor(state_cost.alpha), beta=torch.tensor(action_cost.alpha),
    )
    arrays = {k: v.detach().numpy().copy() for k, v in arrays.items()}
    return net, model, cost, s0, arrays


class _Spec:  # what EnvWrapper._sample_action reads from a dm_env BoundedArray
    def __init__(self, act, lo=-1.0, hi=1.0):
        self.minimum = np.full(act, lo)
        self.maximum = np.full(act, hi)
        self.shape = (act,)


def make_rs(name, obs, act, hidden, n, horizon, keep_states_of=None):
    from src.mbrl.planners import RandomShootingPlanner
    from src.mbrl.env_wrappers import EnvWrapper
    from src.mbrl.agents import MPCPolicy

    net, model, cost, s0, arrays = _reference_problem(obs, act, hidden)
    np.random.seed(1234)
    recorded = EnvWrapper._sample_action(_Spec(act), batch_size=n * horizon)  # reference sampler

    def injected(batch_size):
        assert batch_size == n * horizon
        return recorded.clone()

    trajs, costs = RandomShootingPlanner._generate_trajectories(
        initial_state=s0, model=model, cost=cost, sample_action=injected, horizon=horizon, num_trajectories=n
    )
    states_hno = torch.stack([t[0] for t in trajs], dim=1).detach()  # [H, N, O]
    plan_s, plan_a = RandomShootingPlanner.plan(s0, model, cost, injected, horizon, None, num_trajectories=n)

    class _PlannerN(RandomShootingPlanner):  # MPCPolicy never forwards kwargs (agents.py:48-55)
        @staticmethod
        def plan(**kw):
            return RandomShootingPlanner.plan(num_trajectories=n, **kw)

    policy = MPCPolicy(model=model, cost=cost, planner=_PlannerN, sample_action=injected, horizon=horizon)
    first_action = policy.get_action(dict(timestep=0, observation=s0)).detach().numpy()

    keep = n if keep_states_of is None else keep_states_of
    arrays.update(
        actions=recorded.numpy(), costs=np.asarray(costs, dtype=np.float32),
        idx=np.int64(np.argmin(costs)),
        states_first=states_hno[:, :keep].numpy().copy(),
        plan_states=plan_s.detach().numpy().copy(), plan_actions=plan_a.detach().numpy().copy(),
        first_action=first_action,
        n=np.int64(n), horizon=np.int64(horizon), lo=np.float32(-1), hi=np.float32(1),
    )
    np.savez_compressed(os.path.join(OUT, name), **arrays)
    print(name, "argmin", int(arrays["idx"]), "min cost", float(np.min(costs)))


def make_gd(name, obs, act, hidden, horizon, iters, stop):
    """The reference's GradientDescentPlanner (src/mbrl/planners.py:28-137: Adam(lr 0.01) on one action
    sequence, backprop through the H-step model rollout, early stop on mean |delta a|) run as shipped on a
    recorded initial action sequence.  Records the initial actions, the returned (states, actions) and the
    number of iterations it took (by counting model calls)."""
    from src.mbrl.planners import GradientDescentPlanner
    from src.mbrl.env_wrappers import EnvWrapper

    net, model, cost, s0, arrays = _reference_problem(obs, act, hidden, seed=3)
    np.random.seed(77)
    init = EnvWrapper._sample_action(_Spec(act), batch_size=horizon)   # [H, A], the reference sampler
    calls = [0]

    def counting_model(states, actions):
        calls[0] += 1
        return model(states, actions)

    def injected(batch_size):
        assert batch_size == horizon
        return init.clone()

    states, actions = GradientDescentPlanner.plan(s0, counting_model, cost, injected, horizon, None,
                                                   num_iterations=iters, stop_condition=stop)
    n_calls = calls[0] - horizon  # _initialise_trajectory rolls the model once (planners.py:88-100)
    assert n_calls % horizon == 0
    st = torch.cat([s.reshape(1, -1) for s in states], 0).detach()
    ac = torch.cat([a.reshape(1, -1) for a in actions], 0).detach()
    with torch.no_grad():
        final_cost = float(torch.sum(cost(st[1:], ac)))
    arrays.update(init_actions=init.numpy(), states=st.numpy().copy(), actions=ac.numpy().copy(),
                  iterations_run=np.int64(n_calls // horizon), iters=np.int64(iters), stop=np.float32(stop),
                  horizon=np.int64(horizon), cost_of_returned=np.float32(final_cost), lo=np.float32(-1), hi=np.float32(1))
    np.savez_compressed(os.path.join(OUT, name), **arrays)
    print(name, "iterations run", n_calls // horizon, "cost of returned pair", final_cost)


def make_rs_reward(name, obs, act, hidden, n, horizon):
    """RewardAgent's wiring (src/mbrl/agents.py:342-366): ModelWithReward behind compose(...,
    itemgetter(0)) as the model and compose(..., itemgetter(1)) as the cost, run through the
    reference's own RandomShootingPlanner with an injected sample."""
    from operator import itemgetter
    from src.mbrl.models import ModelWithReward
    from src.mbrl.data import TransitionsDataset
    from src.mbrl.agents import compose
    from src.mbrl.planners import RandomShootingPlanner
    from src.mbrl.env_wrappers import EnvWrapper

    torch.manual_seed(5)
    net = ModelWithReward(obs, act, hidden_units=hidden)
    g = torch.Generator().manual_seed(23)
    stats = {
        "observations": {"mean": torch.randn(obs, generator=g), "std": torch.rand(obs, generator=g) + 0.5},
        "actions": {"mean": 0.1 * torch.randn(act, generator=g), "std": torch.rand(act, generator=g) * 0.3 + 0.4},
        "rewards": {"mean": torch.tensor([0.7]), "std": torch.tensor([1.9])},
    }
    wired = partial(
        net,
        normalize_state=partial(TransitionsDataset.normalize_field, field_name="observations", stats=stats),
        normalize_action=partial(TransitionsDataset.normalize_field, field_name="actions", stats=stats),
        unnormalize_state=partial(TransitionsDataset.unnormalize_field, field_name="observations", stats=stats),
        unnormalize_reward=partial(TransitionsDataset.unnormalize_field, field_name="rewards", stats=stats),
    )
    model, cost = compose(wired, itemgetter(0)), compose(wired, itemgetter(1))
    s0 = stats["observations"]["mean"] + stats["observations"]["std"] * torch.randn(obs, generator=g)
    np.random.seed(4321)
    recorded = EnvWrapper._sample_action(_Spec(act), batch_size=n * horizon)

    def injected(batch_size):
        return recorded.clone()

    trajs, costs = RandomShootingPlanner._generate_trajectories(
        initial_state=s0, model=model, cost=cost, sample_action=injected, horizon=horizon, num_trajectories=n
    )
    plan_s, plan_a = RandomShootingPlanner.plan(s0, model, cost, injected, horizon, None, num_trajectories=n)
    arrays = dict(
        W1=net.linear1.weight, b1=net.linear1.bias, W2=net.linear2.weight, b2=net.linear2.bias,
        W3=net.linear3.weight, b3=net.linear3.bias, W4=net.linear4.weight, b4=net.linear4.bias,
        mu_s=stats["observations"]["mean"], sd_s=stats["observations"]["std"],
        mu_a=stats["actions"]["mean"], sd_a=stats["actions"]["std"],
        mu_r=stats["rewards"]["mean"], sd_r=stats["rewards"]["std"], s0=s0,
    )
    arrays = {k: v.detach().numpy().copy() for k, v in arrays.items()}
    arrays.update(
        actions=recorded.numpy(), costs=np.asarray(costs, dtype=np.float32), idx=np.int64(np.argmin(costs)),
        plan_states=plan_s.detach().numpy().copy(), plan_actions=plan_a.detach().numpy().copy(),
        n=np.int64(n), horizon=np.int64(horizon), lo=np.float32(-1), hi=np.float32(1),
    )
    np.savez_compressed(os.path.join(OUT, name), **arrays)
    print(name, "argmin", int(arrays["idx"]), "min cost", float(np.min(costs)))


def make_rs_linear(name, obs, act, n, horizon):
    """LinearModel (src/mbrl/models.py:113-122, `--model lin`) wired like GoalStateAgent and run through
    the reference's own RandomShootingPlanner with an injected sample."""
    from src.mbrl.models import LinearModel, SmoothAbsLoss, CoshLoss
    from src.mbrl.data import TransitionsDataset
    from src.mbrl.agents import state_action_cost
    from src.mbrl.planners import RandomShootingPlanner
    from src.mbrl.env_wrappers import EnvWrapper

    torch.manual_seed(11)
    net = LinearModel(obs, act)
    g = torch.Generator().manual_seed(31)
    stats = {
        "observations": {"mean": torch.randn(obs, generator=g), "std": torch.rand(obs, generator=g) + 0.5},
        "actions": {"mean": 0.1 * torch.randn(act, generator=g), "std": torch.rand(act, generator=g) * 0.3 + 0.4},
    }
    weights = torch.rand(obs, generator=g) + 0.5
    goal = 0.3 * torch.randn(obs, generator=g)
    state_cost, action_cost = SmoothAbsLoss(weights=weights, goal_state=goal), CoshLoss()
    model = partial(
        net,
        normalize_state=partial(TransitionsDataset.normalize_field, field_name="observations", stats=stats),
        normalize_action=partial(TransitionsDataset.normalize_field, field_name="actions", stats=stats),
        unnormalize_state=partial(TransitionsDataset.unnormalize_field, field_name="observations", stats=stats),
    )
    cost = partial(state_action_cost, state_cost=state_cost, action_cost=action_cost)
    s0 = stats["observations"]["mean"] + stats["observations"]["std"] * torch.randn(obs, generator=g)
    np.random.seed(987)
    recorded = EnvWrapper._sample_action(_Spec(act), batch_size=n * horizon)

    def injected(batch_size):
        return recorded.clone()

    trajs, costs = RandomShootingPlanner._generate_trajectories(
        initial_state=s0, model=model, cost=cost, sample_action=injected, horizon=horizon, num_trajectories=n
    )
    plan_s, plan_a = RandomShootingPlanner.plan(s0, model, cost, injected, horizon, None, num_trajectories=n)
    arrays = dict(
        W1=net.linear1.weight, b1=net.linear1.bias,
        mu_s=stats["observations"]["mean"], sd_s=stats["observations"]["std"],
        mu_a=stats["actions"]["mean"], sd_a=stats["actions"]["std"], cost_w=weights, goal=goal, s0=s0,
        alpha=torch.tensor(state_cost.alpha), beta=torch.tensor(action_cost.alpha),
    )
    arrays = {k: v.detach().numpy().copy() for k, v in arrays.items()}
    arrays.update(
        actions=recorded.numpy(), costs=np.asarray(costs, dtype=np.float32), idx=np.int64(np.argmin(costs)),
        plan_states=plan_s.detach().numpy().copy(), plan_actions=plan_a.detach().numpy().copy(),
        n=np.int64(n), horizon=np.int64(horizon), lo=np.float32(-1), hi=np.float32(1),
    )
    np.savez_compressed(os.path.join(OUT, name), **arrays)
    print(name, "argmin", int(arrays["idx"]), "min cost", float(np.min(costs)))


def make_cem(name, obs, act, hidden, n, horizon, iters, k):
    """Reference-composed CEM: the reference's _generate_trajectories per
    iteration with an injected Gaussian sampler; argsort/mean/std are the only lines that
    are not reference code."""
    from src.mbrl.planners import RandomShootingPlanner

    net, model, cost, s0, arrays = _reference_problem(obs, act, hidden, seed=3)
    g = torch.Generator().manual_seed(99)
    noise = torch.randn(iters, horizon * n, act, generator=g)
    lo, hi = -1.0, 1.0
    mu = torch.zeros(horizon, act)
    sd = torch.ones(horizon, act)
    out = {}
    best = (np.inf, -1, -1)
    for it in range(iters):
        def gauss(batch_size, it=it, mu=mu, sd=sd):
            return torch.clamp(
                mu.repeat_interleave(n, 0) + sd.repeat_interleave(n, 0) * noise[it], min=lo, max=hi
            )

        trajs, costs = RandomShootingPlanner._generate_trajectories(
            initial_state=s0, model=model, cost=cost, sample_action=gauss, horizon=horizon, num_trajectories=n
        )
        elite = np.argsort(costs, kind="stable")[:k]
        acts = torch.stack([trajs[i][1] for i in elite], dim=1)  # [H, k, A]
        if costs[elite[0]] < best[0]:
            best = (costs[elite[0]], it, int(elite[0]))
            out["best_states"] = trajs[elite[0]][0].detach().numpy().copy()
            out["best_actions"] = trajs[elite[0]][1].detach().numpy().copy()
        mu, sd = acts.mean(1), acts.std(1, unbiased=False)
        out[f"costs_{it}"] = np.asarray(costs, dtype=np.float32)
        out[f"elite_{it}"] = elite.astype(np.int64)
        out[f"mu_{it}"] = mu.numpy().copy()
        out[f"sd_{it}"] = sd.numpy().copy()
    arrays.update(out)
    arrays.update(
        noise=noise.numpy(), n=np.int64(n), horizon=np.int64(horizon), iters=np.int64(iters), k=np.int64(k),
        best_cost=np.float32(best[0]), best_it=np.int64(best[1]), best_idx=np.int64(best[2]),
        lo=np.float32(lo), hi=np.float32(hi),
    )
    np.savez_compressed(os.path.join(OUT, name), **arrays)
    print(name, "best", best)


def make_ring_world():
    """src/mbrl/test_random_shooting.py:6-25 with all 3^5 action sequences injected:
    the reference planner's answer on the enumerated population (known optimum 17)."""
    import itertools
    from src.mbrl.planners import RandomShootingPlanner

    world, goal, hz = 10, torch.tensor(9, dtype=torch.float), 5
    s0 = torch.tensor([2], dtype=torch.float)
    seqs = torch.tensor(list(itertools.product([-1.0, 0.0, 1.0], repeat=hz)))  # [243, 5]
    n = seqs.shape[0]
    acts = seqs.t().contiguous().view(hz * n, 1)

    def model(s, a):
        return torch.fmod(torch.fmod(s + a, world) + world, world)

    def cost(s, a):
        return torch.abs(s - goal)

    trajs, costs = RandomShootingPlanner._generate_trajectories(s0, model, cost, lambda batch_size: acts, hz, n)
    ps, pa = RandomShootingPlanner.plan(s0, model, cost, lambda batch_size: acts, hz, None, num_trajectories=n)
    np.savez_compressed(
        os.path.join(OUT, "ring_world.npz"), actions=acts.numpy(), costs=np.asarray(costs, dtype=np.float32),
        plan_states=ps.numpy().copy(), plan_actions=pa.numpy().copy(),
    )
    print("ring_world min cost", costs.min(), "plan", ps.flatten().tolist(), pa.flatten().tolist())


def make_tolerance():
    """Known-answer grid for dm_control.utils.rewards.tolerance
    (dm_control/dm_control/utils/rewards.py:88-130), the building block of every suite
    task reward."""
    sys.path.insert(0, os.path.join(REF, "dm_control"))
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "_ref_rewards", os.path.join(REF, "dm_control", "dm_control", "utils", "rewards.py")
    )
    rewards = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rewards)
    x = np.linspace(-4.0, 4.0, 161)
    cases = []
    vals = []
    for sig_id, sig in enumerate(["gaussian", "linear", "quadratic", "hyperbolic", "long_tail", "cosine", "tanh_squared"]):
        for (lo, hi, margin, vam) in [(0.0, 0.0, 1.0, 0.1), (-0.5, 1.0, 2.0, 0.1), (1.0, np.inf, 0.5, 0.5),
                                      (1.2, np.inf, 0.3, 0.1), (-2.0, 2.0, 0.0, 0.1)]:
            if sig in ("cosine", "linear", "quadratic") and not (0 <= vam < 1):
                continue
            cases.append([sig_id, lo, hi, margin, vam])
            vals.append(rewards.tolerance(x, bounds=(lo, hi), margin=margin, sigmoid=sig, value_at_margin=vam))
    np.savez_compressed(os.path.join(OUT, "tolerance.npz"), x=x, cases=np.array(cases), values=np.array(vals))
    print("tolerance cases", len(cases))

    # Humanoid.get_reward (dm_control/suite/humanoid.py:187-211) composed with the reference's own
    # tolerance() on synthetic physics quantities (MuJoCo itself is not available): pins the
    # restated humanoid task cost (oracle/task_costs.humanoid_cost).
    rng = np.random.default_rng(7)
    m = 512
    head = rng.uniform(0.2, 1.8, m); zz = rng.uniform(-1.0, 1.0, m)
    com = rng.normal(size=(m, 3)) * np.array([6.0, 3.0, 1.0]); ctrl = rng.uniform(-1.3, 1.3, size=(m, 21))
    rew = np.empty(m)
    for i in range(m):
        standing = rewards.tolerance(head[i], bounds=(1.4, float("inf")), margin=1.4 / 4)
        upright = rewards.tolerance(zz[i], bounds=(0.9, float("inf")), sigmoid="linear", margin=1.9, value_at_margin=0)
        small_control = rewards.tolerance(ctrl[i], margin=1, value_at_margin=0, sigmoid="quadratic").mean()
        small_control = (4 + small_control) / 5
        com_velocity = np.linalg.norm(com[i][[0, 1]])
        move = rewards.tolerance(com_velocity, bounds=(10, float("inf")), margin=10, value_at_margin=0, sigmoid="linear")
        move = (5 * move + 1) / 6
        rew[i] = small_control * standing * upright * move
    np.savez_compressed(os.path.join(OUT, "humanoid_reward.npz"), head_height=head, torso_upright=zz, com_velocity=com,
                        control=ctrl, reward=rew)
    print("humanoid reward samples", m, "mean", rew.mean())

    # Cheetah.get_reward (dm_control/suite/cheetah.py:91-97) and PlanarWalker.get_reward at move_speed 1
    # (walker-walk, dm_control/suite/walker.py:135-158) composed with the reference's own tolerance() on
    # synthetic physics quantities: pins oracle/task_costs.cheetah_run_cost / walker_walk_cost.
    rng = np.random.default_rng(11)
    speed = rng.uniform(-4.0, 14.0, m)
    rew_c = np.array([rewards.tolerance(v, bounds=(10, float("inf")), margin=10, value_at_margin=0, sigmoid="linear") for v in speed])
    height = rng.uniform(0.2, 1.6, m); zz = rng.uniform(-1.0, 1.0, m); vel = rng.uniform(-1.5, 2.5, m)
    rew_w = np.empty(m)
    for i in range(m):
        standing = rewards.tolerance(height[i], bounds=(1.2, float("inf")), margin=1.2 / 2)
        upright = (1 + zz[i]) / 2
        stand_reward = (3 * standing + upright) / 4
        move_reward = rewards.tolerance(vel[i], bounds=(1, float("inf")), margin=1 / 2, value_at_margin=0.5, sigmoid="linear")
        rew_w[i] = stand_reward * (5 * move_reward + 1) / 6
    np.savez_compressed(os.path.join(OUT, "locomotion_reward.npz"), cheetah_speed=speed, cheetah_reward=rew_c,
                        walker_height=height, walker_upright=zz, walker_velocity=vel, walker_reward=rew_w)
    print("cheetah / walker reward samples", m, "means", rew_c.mean(), rew_w.mean())


if __name__ == "__main__":
    _install_stubs()
    make_ring_world()
    make_rs("rs_cartpole.npz", obs=5, act=1, hidden=50, n=1000, horizon=20)
    make_rs("rs_cheetah_small.npz", obs=17, act=6, hidden=200, n=256, horizon=30, keep_states_of=32)
    make_cem("cem_cheetah_small.npz", obs=17, act=6, hidden=200, n=256, horizon=30, iters=3, k=25)
    make_cem("cem_cartpole_small.npz", obs=5, act=1, hidden=50, n=512, horizon=30, iters=4, k=51)
    make_tolerance()
    make_rs_reward("rs_reward_head.npz", obs=17, act=6, hidden=50, n=256, horizon=20)
    make_rs_linear("rs_linear_model.npz", obs=9, act=3, n=200, horizon=12)
    make_gd("gd_small.npz", obs=9, act=3, hidden=64, horizon=12, iters=25, stop=0.002)
    make_gd("gd_early_stop.npz", obs=9, act=3, hidden=64, horizon=12, iters=40, stop=0.0095)
    make_gd("gd_cheetah.npz", obs=17, act=6, hidden=200, horizon=30, iters=40, stop=0.002)
