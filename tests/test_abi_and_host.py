"""CPU-side checks: the C-ABI library builds/loads and exports every symbol the header
declares (no compute calls -- there is no GPU here), the host adaptor recognises the
reference's partials, and misuse fails loudly instead of falling back."""
import ctypes
import functools
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from mbrl_b200 import native
    lib = native.load_library()
    header = open(os.path.join(ROOT, "include", "mbrl_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char\*)\s+(mbrl_\w+)\s*\(", header, re.M))
    assert declared == set(native.ABI_SYMBOLS), declared ^ set(native.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mbrl_abi_version() == 2


def test_struct_layouts_match_header():
    from mbrl_b200 import native
    assert ctypes.sizeof(native.MbrlConfig) == 40
    assert ctypes.sizeof(native.MbrlPlanInfo) == 16 == native.PLAN_INFO_DTYPE.itemsize
    assert ctypes.sizeof(native.MbrlPlanArgs) == 72


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from mbrl_b200 import native
    with pytest.raises(native.MbrlError):
        native.NativePlanner(5, 1, 50, 20, 1000)


def test_invalid_arguments_are_rejected_before_cuda():
    from mbrl_b200 import native
    lib = native.load_library()
    cfg = native.MbrlConfig(0, 1, 50, 20, 1000, 1, 1, 1, 0, 0)
    handle = ctypes.c_void_p()
    assert lib.mbrl_create(ctypes.byref(cfg), ctypes.byref(handle)) == -1
    assert b"obs_dim" in lib.mbrl_last_error()
    assert lib.mbrl_topk(None, 1, 1, 1, None, None, None, None) == -1


def test_adaptor_introspects_reference_partials():
    from mbrl_b200.adaptor import PlanningProblem, problem_from_callables

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.linear1, self.linear2, self.linear3 = torch.nn.Linear(7, 12), torch.nn.Linear(12, 12), torch.nn.Linear(12, 5)
            self.noise = None

    def norm(field_value, field_name, stats):
        return field_value

    stats = {"observations": {"mean": torch.zeros(5), "std": torch.ones(5)},
             "actions": {"mean": torch.zeros(2), "std": torch.ones(2)}}
    net = Net()
    model = functools.partial(net, normalize_state=functools.partial(norm, field_name="observations", stats=stats),
                              normalize_action=functools.partial(norm, field_name="actions", stats=stats),
                              unnormalize_state=functools.partial(norm, field_name="observations", stats=stats))
    sc = type("SC", (), dict(weights=torch.ones(5), goal_state=torch.zeros(5), alpha=0.4))()
    ac = type("AC", (), dict(alpha=0.25))()
    cost = functools.partial(lambda s, a, state_cost, action_cost: 0, state_cost=sc, action_cost=ac)
    spec = type("Spec", (), dict(minimum=np.array([-5.0, -1.0]), maximum=np.array([0.5, 1.0]), shape=(2,)))()
    sampler = functools.partial(lambda action_spec, batch_size=None: None, action_spec=spec)
    prob, fp = problem_from_callables(model, cost, sampler)
    assert (prob.obs_dim, prob.act_dim, prob.hidden) == (5, 2, 12)
    assert (prob.act_lo, prob.act_hi) == (-3.0, 0.5)  # dim-0 bounds, clipped to +-3 (env_wrappers.py:52-55)
    _, fp_same = problem_from_callables(model, cost, sampler)
    assert fp == fp_same
    with torch.no_grad():
        net.linear2.weight.mul_(0.5)  # in-place training step
    assert problem_from_callables(model, cost, sampler)[1] != fp
    stats["observations"] = {"mean": torch.ones(5), "std": torch.ones(5)}  # add_rollouts replaces entries
    assert problem_from_callables(model, cost, sampler)[1][1] != fp[1]
    with pytest.raises(TypeError):
        problem_from_callables(lambda s, a: s, cost, sampler)
    with pytest.raises(TypeError):
        problem_from_callables(model, lambda s, a: 0, sampler)
    assert isinstance(problem_from_callables(prob, None, None)[0], PlanningProblem)


def test_planner_classes_pickle():
    import pickle
    import mbrl_b200
    for cls in (mbrl_b200.RandomShootingPlanner, mbrl_b200.CEMPlanner):
        assert pickle.loads(pickle.dumps(cls)) is cls
        assert isinstance(cls.__dict__["plan"], staticmethod)


def test_synthetic_inputs_identical_for_product_and_oracle():
    from mbrl_b200.synthetic import synthetic_problem, synthetic_state
    from oracle import planner_oracle as po
    a, b = synthetic_problem(17, 6, 200), po.synthetic_params(17, 6, 200)
    for name in ("W1", "b1", "W2", "b2", "W3", "b3", "mu_s", "sd_s", "mu_a", "sd_a", "cost_w", "goal"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert torch.equal(synthetic_state(a, 3), po.synthetic_state(b, 3))


def test_cem_warm_start_modes_host_logic(monkeypatch):
    """CEMPlanner's warm start is host logic over the C ABI: checked with a
    recording stand-in for the native handle (no GPU)."""
    import numpy as np
    import pytest
    import torch
    from mbrl_b200 import native, planners
    from mbrl_b200.adaptor import PlanningProblem

    H, A, O = 4, 2, 3
    calls = []

    class FakeHandle:
        def plan(self, s0, iters, k, mode, seed, injected=None, mu0=None, sd0=None, return_mean=False,
                 want_dist=False, actions_only=False, warm_start=0, warm_std=0.0):
            calls.append(dict(mu0=None if mu0 is None else np.array(mu0), sd0=None if sd0 is None else np.array(sd0),
                              want_dist=want_dist, warm_start=warm_start, warm_std=warm_std))
            return dict(states=np.zeros((1, H, O), np.float32), actions=np.zeros((1, H, A), np.float32),
                        mu=None, sd=None, info=None)

    ent = dict(handle=FakeHandle(), calls=0)

    class Prob:
        act_dim, act_lo, act_hi = A, -1.0, 1.0

    monkeypatch.setattr(planners, "_get_handle", lambda *a, **kw: (ent, Prob))
    plan = planners.CEMPlanner.plan
    prev = (torch.zeros(H - 1, O), torch.full((H, A), 0.25))
    # shift_mean keeps the mean RESIDENT ON THE DEVICE: nothing is uploaded or read back; the first step of an
    # episode (initial_trajectory None) only asks the handle to keep its final mean, later steps also use it
    plan(torch.zeros(O), None, None, None, H, None, warm_start="shift_mean", num_trajectories=8)
    assert calls[-1]["mu0"] is None and not calls[-1]["want_dist"] and calls[-1]["warm_start"] == native.WARM_KEEP
    plan(torch.zeros(O), None, None, None, H, prev, warm_start="shift_mean", num_trajectories=8, init_std=0.3)
    assert calls[-1]["mu0"] is None and calls[-1]["warm_start"] == native.WARM_KEEP | native.WARM_USE
    assert calls[-1]["warm_std"] == pytest.approx(0.3)
    plan(torch.zeros(O), None, None, None, H, None, warm_start="shift_mean", num_trajectories=8)
    assert calls[-1]["warm_start"] == native.WARM_KEEP  # episode start drops the remembered mean
    # trajectory (default): the reference's handed-over action sequence seeds the mean
    plan(torch.zeros(O), None, None, None, H, prev, num_trajectories=8)
    np.testing.assert_array_equal(calls[-1]["mu0"], np.full((H, A), 0.25, np.float32))
    np.testing.assert_array_equal(calls[-1]["sd0"], np.ones((H, A), np.float32))
    # none: never warm-started
    plan(torch.zeros(O), None, None, None, H, prev, warm_start="none", num_trajectories=8)
    assert calls[-1]["mu0"] is None and calls[-1]["warm_start"] == 0
    with pytest.raises(ValueError):
        plan(torch.zeros(O), None, None, None, H, prev, warm_start="bogus", num_trajectories=8)


def _reward_agent_wiring(obs=4, act=2, hidden=8, seed=0):
    """The callables RewardAgent hands to MPCPolicy (src/mbrl/agents.py:342-366), rebuilt with
    stand-ins shaped like the reference's ModelWithReward / compose / normalize_field."""
    from functools import partial
    from operator import itemgetter
    import torch

    class ModelWithReward(torch.nn.Module):  # src/mbrl/models.py:125-163
        def __init__(self):
            super().__init__()
            self.linear1 = torch.nn.Linear(obs + act, hidden)
            self.linear2 = torch.nn.Linear(hidden, hidden)
            self.linear3 = torch.nn.Linear(hidden, obs)
            self.linear4 = torch.nn.Linear(hidden, 1)

        def forward(self, state, action, normalize_state=None, unnormalize_state=None, normalize_action=None,
                    unnormalize_reward=None):
            x = torch.cat([normalize_state(state), normalize_action(action)], dim=1)
            x = torch.relu(self.linear2(torch.relu(self.linear1(x))))
            return unnormalize_state(self.linear3(x)), unnormalize_reward(self.linear4(x))

    def normalize_field(x, field_name, stats):
        return (x - stats[field_name]["mean"]) / stats[field_name]["std"]

    def unnormalize_field(x, field_name, stats):
        return x * stats[field_name]["std"] + stats[field_name]["mean"]

    def compose(a, b):  # src/mbrl/agents.py:290-295
        def ab(*args, **kwargs):
            return b(a(*args, **kwargs))
        return ab

    torch.manual_seed(seed)
    net = ModelWithReward()
    stats = {"observations": {"mean": torch.randn(obs), "std": torch.rand(obs) + 0.5},
             "actions": {"mean": torch.zeros(act), "std": torch.ones(act) * 0.6},
             "rewards": {"mean": torch.tensor([0.3]), "std": torch.tensor([2.0])}}
    wired = partial(net, normalize_state=partial(normalize_field, field_name="observations", stats=stats),
                    normalize_action=partial(normalize_field, field_name="actions", stats=stats),
                    unnormalize_state=partial(unnormalize_field, field_name="observations", stats=stats),
                    unnormalize_reward=partial(unnormalize_field, field_name="rewards", stats=stats))
    return net, stats, compose(wired, itemgetter(0)), compose(wired, itemgetter(1)), wired


def test_adaptor_recognises_reward_agent_wiring():
    import pytest
    from mbrl_b200 import native
    from mbrl_b200.adaptor import problem_from_callables
    net, stats, model, cost, wired = _reward_agent_wiring()
    prob, fp = problem_from_callables(model, cost, None)
    assert prob.cost_kind == native.COST_REWARD_HEAD
    assert prob.W4 is net.linear4.weight and prob.b4 is net.linear4.bias
    assert prob.mu_r is stats["rewards"]["mean"] and prob.sd_r is stats["rewards"]["std"]
    assert (prob.obs_dim, prob.act_dim, prob.hidden) == (4, 2, 8)
    with torch.no_grad():
        net.linear4.weight.add_(1.0)  # retraining in place changes the fingerprint (models.py:84-86)
    assert problem_from_callables(model, cost, None)[1] != fp
    with pytest.raises(TypeError):  # the bare partial is not how RewardAgent passes it
        problem_from_callables(wired, cost, None)
    with pytest.raises(TypeError):  # swapped itemgetters
        problem_from_callables(cost, model, None)


def test_linear_model_embedding_is_exact():
    """LinearModel runs through an exact ReLU embedding (adaptor.embed_linear_model): the embedded
    3-layer MLP reproduces W x + b to fp32 rounding of the summation order only."""
    from functools import partial
    from mbrl_b200.adaptor import embed_linear_model, problem_from_callables
    torch.manual_seed(3)

    class LinearModel(torch.nn.Module):  # src/mbrl/models.py:113-122
        def __init__(self):
            super().__init__()
            self.linear1 = torch.nn.Linear(12, 9)
            self.noise = None
    net = LinearModel()
    W1, b1, W2, b2, W3, b3 = (torch.from_numpy(a) for a in embed_linear_model(net.linear1.weight, net.linear1.bias))
    x = torch.randn(64, 12) * 3
    h = torch.relu(torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(x, W1, b1)), W2, b2))
    y = torch.nn.functional.linear(h, W3, b3)
    torch.testing.assert_close(y, net.linear1(x).detach(), rtol=1e-6, atol=1e-6)

    class Cost:
        weights, goal_state, alpha = torch.ones(9), torch.zeros(9), 0.4
    class ACost:
        alpha = 0.25
    stats = {"observations": {"mean": torch.zeros(9), "std": torch.ones(9)}, "actions": {"mean": torch.zeros(3), "std": torch.ones(3)}}
    f = lambda x, field_name, stats: x
    model = partial(net, normalize_state=partial(f, field_name="observations", stats=stats),
                    normalize_action=partial(f, field_name="actions", stats=stats),
                    unnormalize_state=partial(f, field_name="observations", stats=stats))
    prob, fp = problem_from_callables(model, partial(f, state_cost=Cost, action_cost=ACost), None)
    assert (prob.obs_dim, prob.act_dim, prob.hidden) == (9, 3, 24)
    with torch.no_grad():
        net.linear1.weight.mul_(2.0)  # retraining in place: new fingerprint, new embedding
    prob2, fp2 = problem_from_callables(model, partial(f, state_cost=Cost, action_cost=ACost), None)
    assert fp2 != fp and not np.array_equal(np.asarray(prob2.W3), np.asarray(prob.W3))
