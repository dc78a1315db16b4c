"""GPU parity tests (run on the B200 with ``-m gpu``).  Every computation goes through the
C ABI (libmbrl_b200.so); the oracle (oracle/) is only the checker."""
import functools

import numpy as np
import pytest
import torch

from mbrl_helpers import load_golden, params_from_golden
from oracle import philox as ophilox
from oracle import planner_oracle as po

pytestmark = pytest.mark.gpu

# fp32 engine tolerances: per-step predicted states and trajectory costs against the fp32
# CPU planner (only the GEMM summation order differs from MKL).
FP32_STATE_RTOL, FP32_STATE_ATOL = 2e-5, 2e-5
FP32_COST_RTOL = 2e-5


@pytest.fixture(scope="module")
def native():
    from mbrl_b200 import native as n
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    n.load_library()
    return n


def _planner(native, p, horizon, n, envs=1, iters=1, engine="fp32", max_elites=None):
    h = native.NativePlanner(p.obs_dim, p.act_dim, p.hidden, horizon, n, envs, iters, max_elites, engine)
    h.set_weights(p.W1, p.b1, p.W2, p.b2, p.W3, p.b3)
    h.set_norm(p.mu_s, p.sd_s, p.mu_a, p.sd_a)
    h.set_cost(p.cost_w, p.goal, p.alpha, p.beta)
    h.set_action_bounds(p.act_lo, p.act_hi)
    return h


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ---------------------------------------------------------------------------------------
# sampler
# ---------------------------------------------------------------------------------------
def test_philox_known_answers_and_random_counters(native):
    rng = np.random.default_rng(0)
    ctr = rng.integers(0, 2 ** 32, size=(4099, 4), dtype=np.uint64).astype(np.uint32)
    key = rng.integers(0, 2 ** 32, size=(4099, 2), dtype=np.uint64).astype(np.uint32)
    ctr[0], key[0] = 0, 0
    ctr[1], key[1] = 0xFFFFFFFF, 0xFFFFFFFF
    ctr[2], key[2] = [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]
    out = native.philox_raw(_cuda(ctr.view(np.int32)), _cuda(key.view(np.int32))).cpu().numpy().view(np.uint32)
    assert out[0].tolist() == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert out[1].tolist() == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert out[2].tolist() == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    np.testing.assert_array_equal(out, ophilox.philox4x32_10(ctr, key))


@pytest.mark.parametrize("act_dim,n,envs", [(6, 300, 1), (1, 257, 2), (21, 64, 3)])
def test_sampler_matches_oracle(native, act_dim, n, envs):
    p = po.synthetic_params(7, act_dim, 16)
    H = 5
    h = _planner(native, p, H, n, envs)
    h.set_action_bounds(-0.8, 0.9)
    seed, it, coff, eoff = 0x1234567890ABCDEF, 3, 1000, 5
    u = h.sample(native.SAMPLE_UNIFORM, seed, it, cand_offset=coff, env_offset=eoff).cpu().numpy()
    ref_u = np.concatenate(
        [ophilox.uniform(seed, it, H, n, act_dim, -0.8, 0.9, coff, eoff + e).reshape(H, n, act_dim) for e in range(envs)],
        axis=1).reshape(H * n * envs, act_dim)
    np.testing.assert_array_equal(u, ref_u)  # integer generator + exact fp32 affine map
    mu = torch.zeros(envs, H, act_dim, device="cuda")
    sd = torch.ones(envs, H, act_dim, device="cuda")
    h.set_action_bounds(-100.0, 100.0)
    z = h.sample(native.SAMPLE_GAUSSIAN, seed, it, mu, sd, coff, eoff).cpu().numpy()
    ref_z = np.concatenate(
        [ophilox.standard_normal(seed, it, H, n, act_dim, coff, eoff + e).reshape(H, n, act_dim) for e in range(envs)],
        axis=1).reshape(H * n * envs, act_dim)
    # Box-Muller runs on the SFU (lg2 / sqrt / sin / cos approx, csrc/philox.cuh): <= 4e-6 from numpy's libm
    np.testing.assert_allclose(z, ref_z, rtol=0, atol=4e-6)
    # clip + affine
    mu2 = torch.rand(envs, H, act_dim, device="cuda") - 0.5
    sd2 = torch.rand(envs, H, act_dim, device="cuda") + 0.1
    h.set_action_bounds(-0.5, 0.5)
    a = h.sample(native.SAMPLE_GAUSSIAN, seed, it, mu2, sd2, coff, eoff).cpu()
    zz = torch.from_numpy(z).view(H, envs, n, act_dim)
    want = torch.clamp(mu2.cpu().permute(1, 0, 2)[:, :, None, :] + sd2.cpu().permute(1, 0, 2)[:, :, None, :] * zz, -0.5, 0.5)
    np.testing.assert_allclose(a.view(H, envs, n, act_dim).numpy(), want.numpy(), rtol=0, atol=1e-5)
    assert a.min() >= -0.5 and a.max() <= 0.5 and (a == 0.5).any() and (a == -0.5).any()


# ---------------------------------------------------------------------------------------
# rollout + cost
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["rs_cartpole.npz", "rs_cheetah_small.npz"])
def test_rollout_matches_reference_golden(native, name):
    g = load_golden(name)
    p = params_from_golden(g)
    n, H = int(g["n"]), int(g["horizon"])
    h = _planner(native, p, H, n)
    costs, states, actions = h.rollout(_cuda(g["s0"][None]), native.SAMPLE_INJECT_ACTIONS, d_injected=_cuda(g["actions"]),
                                       want_states=True, want_actions=True)
    np.testing.assert_array_equal(actions.cpu().numpy(), g["actions"])
    np.testing.assert_allclose(costs.cpu().numpy(), g["costs"], rtol=FP32_COST_RTOL)
    keep = g["states_first"].shape[1]
    got = states.cpu().numpy().reshape(H, n, -1)[:, :keep]
    np.testing.assert_allclose(got, g["states_first"], rtol=FP32_STATE_RTOL, atol=FP32_STATE_ATOL)
    assert int(torch.argmin(costs)) == int(g["idx"])


@pytest.mark.parametrize("name", ["rs_cartpole.npz", "rs_cheetah_small.npz"])
def test_rs_plan_matches_reference_golden(native, name):
    g = load_golden(name)
    p = params_from_golden(g)
    n, H = int(g["n"]), int(g["horizon"])
    h = _planner(native, p, H, n)
    out = h.plan(g["s0"], 1, 1, native.SAMPLE_INJECT_ACTIONS, injected=g["actions"])
    assert int(out["info"]["best_index"][0]) == int(g["idx"])
    np.testing.assert_allclose(out["info"]["best_cost"][0], g["costs"].min(), rtol=FP32_COST_RTOL)
    np.testing.assert_array_equal(out["actions"][0], g["plan_actions"])
    np.testing.assert_array_equal(out["actions"][0][0], g["first_action"])
    np.testing.assert_allclose(out["states"][0], g["plan_states"], rtol=FP32_STATE_RTOL, atol=FP32_STATE_ATOL)


def test_actions_only_plan_skips_replay(native):
    g = load_golden("rs_cartpole.npz")
    p = params_from_golden(g)
    n, H = int(g["n"]), int(g["horizon"])
    h = _planner(native, p, H, n)
    out = h.plan(g["s0"], 1, 1, native.SAMPLE_INJECT_ACTIONS, injected=g["actions"], actions_only=True)
    np.testing.assert_array_equal(out["actions"][0], g["plan_actions"])
    assert int(out["info"]["best_index"][0]) == int(g["idx"]) and not out["states"].any()


def test_dmc_cartpole_task_cost_epilogue(native):
    """Cost option C: 1 - cartpole swing-up reward restated on observations,
    against the oracle built on the reference's rewards.tolerance (fp32 engine)."""
    from oracle import task_costs
    g = load_golden("rs_cartpole.npz")
    p = params_from_golden(g)
    n, H = int(g["n"]), int(g["horizon"])
    h = _planner(native, p, H, n)
    h.set_cost(kind=native.COST_DMC_CARTPOLE_SWINGUP)
    acts = torch.from_numpy(g["actions"])
    costs, states, _ = h.rollout(_cuda(g["s0"][None]), native.SAMPLE_INJECT_ACTIONS, d_injected=acts.cuda(), want_states=True)
    st, _ = po.rollout_costs(p, torch.from_numpy(g["s0"]), acts, H, n)
    want = task_costs.cartpole_swingup_cost(st.numpy(), g["actions"]).reshape(H, n).sum(0)
    np.testing.assert_allclose(costs.cpu().numpy(), want, rtol=2e-5, atol=2e-5)


TASK_COST_TOL = {"fp32": 1e-4, "fp16": 3e-3, "bf16": 2e-2}  # abs + rel, on sums of H per-step costs in [0, 1]


@pytest.mark.parametrize("engine", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("task", ["cartpole", "cheetah", "walker", "humanoid"])
def test_dm_control_task_costs_on_every_engine(native, engine, task):
    """The dm_control task-cost epilogues on the fp32 engine AND the tensor-core engines:
    cartpole-swingup (cartpole.py:216-226), cheetah-run (cheetah.py:91-97, speed proxy obs[8]), walker-walk
    (walker.py:135-158, speed proxy obs[16]), humanoid-run (humanoid.py:187-211), each against the oracle
    that is pinned to the reference's own rewards.tolerance (tests/test_task_costs_oracle.py).  State
    statistics put the picked entries in the rewards' active range.  Tolerance: TASK_COST_TOL (16-bit
    operand rounding of ~5e-4 of the state scale passes through tolerance() slopes of O(1/margin))."""
    from oracle import task_costs
    O, A, U = {"cartpole": (5, 1, 50), "cheetah": (17, 6, 200), "walker": (24, 6, 200), "humanoid": (67, 21, 128)}[task]
    p = po.synthetic_params(O, A, U, seed=4)
    stats = {"cartpole": {0: (0.0, 1.5), 1: (0.2, 0.5), 4: (0.0, 4.0)},
             "cheetah": {8: (5.0, 4.0)},
             "walker": {14: (1.0, 0.4), 0: (0.3, 0.5), 16: (0.7, 0.6)},
             "humanoid": {21: (1.2, 0.4), 36: (0.5, 0.6), 37: (4.0, 5.0), 38: (0.0, 3.0)}}[task]
    for i, (mu_i, sd_i) in stats.items():
        p.mu_s[i], p.sd_s[i] = mu_i, sd_i
    kind = {"cartpole": native.COST_DMC_CARTPOLE_SWINGUP, "cheetah": native.COST_DMC_CHEETAH_RUN,
            "walker": native.COST_DMC_WALKER_WALK, "humanoid": native.COST_DMC_HUMANOID_RUN}[task]
    H, n = 8, 300
    h = _planner(native, p, H, n, engine=engine)
    h.set_cost(kind=kind)
    g = torch.Generator().manual_seed(2)
    s0 = po.synthetic_state(p, 3)
    acts = torch.rand(H * n, A, generator=g) * 2.6 - 1.3  # beyond +-1: the quadratic control tolerance saturates
    h.set_action_bounds(-1.3, 1.3)
    costs, states, _ = h.rollout(s0[None].cuda(), native.SAMPLE_INJECT_ACTIONS, d_injected=acts.cuda(), want_states=True)
    st, _ = po.rollout_costs(p, s0, acts, H, n)
    fn = {"cartpole": lambda: task_costs.cartpole_swingup_cost(st.numpy(), acts.numpy()),
          "cheetah": lambda: task_costs.cheetah_run_cost(st.numpy()),
          "walker": lambda: task_costs.walker_walk_cost(st.numpy()),
          "humanoid": lambda: task_costs.humanoid_cost(st.numpy(), acts.numpy())}[task]
    want = fn().reshape(H, n).sum(0)
    assert want.std() > 0.01, "the test must exercise the reward's active range"
    tol = TASK_COST_TOL[engine]
    err = np.abs(costs.cpu().numpy() - want).max()
    print(f"task cost {task} on {engine}: max abs err {err:.2e} (cost range {want.min():.3f}..{want.max():.3f})")
    np.testing.assert_allclose(costs.cpu().numpy(), want, rtol=tol, atol=tol)
    # and a whole plan minimises it: the reported best cost is the task cost of the emitted sequence
    hp = _planner(native, p, H, 512, 1, 3, engine=engine)
    hp.set_cost(kind=kind)
    out = hp.plan(s0.numpy(), 3, 51, native.SAMPLE_GAUSSIAN, seed=5)
    a_best = torch.from_numpy(out["actions"][0])
    st_b, _ = po.rollout_costs(p, s0, a_best, H, 1)
    fnb = {"cartpole": lambda: task_costs.cartpole_swingup_cost(st_b.numpy(), a_best.numpy()),
           "cheetah": lambda: task_costs.cheetah_run_cost(st_b.numpy()),
           "walker": lambda: task_costs.walker_walk_cost(st_b.numpy()),
           "humanoid": lambda: task_costs.humanoid_cost(st_b.numpy(), a_best.numpy())}[task]
    np.testing.assert_allclose(out["info"]["best_cost"][0], fnb().sum(), rtol=tol, atol=tol)


def test_rollout_ragged_rows_and_batched_envs(native):
    """N not a multiple of the row tile, several environments with distinct s0."""
    p = po.synthetic_params(9, 3, 40, seed=5)
    H, n, E = 7, 77, 3
    h = _planner(native, p, H, n, E)
    g = torch.Generator().manual_seed(1)
    s0 = p.mu_s + p.sd_s * torch.randn(E, 9, generator=g)
    acts = torch.rand(H * E * n, 3, generator=g) * 2 - 1  # rows r = env*n + cand, step-major
    costs, states, _ = h.rollout(s0.cuda(), native.SAMPLE_INJECT_ACTIONS, d_injected=acts.cuda(), want_states=True)
    a4 = acts.view(H, E, n, 3)
    for e in range(E):
        st, c = po.rollout_costs(p, s0[e], a4[:, e].reshape(H * n, 3), H, n)
        np.testing.assert_allclose(costs.cpu().numpy()[e * n:(e + 1) * n], c, rtol=FP32_COST_RTOL)
        np.testing.assert_allclose(states.cpu().view(H, E, n, 9)[:, e].numpy(), st.view(H, n, 9).numpy(),
                                   rtol=FP32_STATE_RTOL, atol=FP32_STATE_ATOL)


def test_humanoid_shape_batched_envs_fp32_engine(native):
    """BASELINE config 5 shape (obs 67, act 21, hidden 512, batched independent environments) on
    the fp32 engine (tests/test_gpu_tc.py runs the same shape on the weight-streaming tcgen05 kernel)."""
    p = po.synthetic_params(67, 21, 512, seed=9)
    H, n, E = 6, 96, 3
    h = _planner(native, p, H, n, E, iters=2)
    g = torch.Generator().manual_seed(4)
    s0 = p.mu_s + p.sd_s * torch.randn(E, 67, generator=g)
    acts = torch.rand(H * E * n, 21, generator=g) * 2 - 1
    costs, states, _ = h.rollout(s0.cuda(), native.SAMPLE_INJECT_ACTIONS, d_injected=acts.cuda(), want_states=True)
    a4 = acts.view(H, E, n, 21)
    for e in range(E):
        st, c = po.rollout_costs(p, s0[e], a4[:, e].reshape(H * n, 21), H, n)
        np.testing.assert_allclose(costs.cpu().numpy()[e * n:(e + 1) * n], c, rtol=5e-5)
        np.testing.assert_allclose(states.cpu().view(H, E, n, 67)[:, e].numpy(), st.view(H, n, 67).numpy(), rtol=1e-4, atol=1e-4)
    out = h.plan(s0.numpy(), 2, 9, native.SAMPLE_GAUSSIAN, seed=1)  # batched CEM: per-environment elites and plans
    assert out["actions"].shape == (E, H, 21) and np.isfinite(out["states"]).all()
    for e in range(E):
        _, c = po.rollout_costs(p, s0[e], torch.from_numpy(out["actions"][e]), H, 1)
        np.testing.assert_allclose(out["info"]["best_cost"][e], c[0], rtol=5e-5)
    with pytest.raises(native.MbrlError, match="hidden"):
        native.NativePlanner(67, 21, 600, H, n, E, engine="fp16")  # > 512 hidden units: no tensor-core kernel


# ---------------------------------------------------------------------------------------
# elite select
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,k,segments", [(1, 1, 1), (5, 5, 1), (1000, 1, 1), (1024, 100, 2), (4096, 409, 1),
                                           (16384, 1638, 1), (2048, 204, 7), (131072, 13107, 1), (3001, 3000, 2)])
def test_topk_bit_exact_with_ties(native, n, k, segments):
    rng = np.random.default_rng(n + k)
    # coarse quantisation -> many exact ties, including at the k-th value
    costs = np.round(rng.normal(size=(segments, n)) * 20).astype(np.float32) / 4
    costs[:, rng.integers(0, n, size=max(1, n // 50))] *= -1
    idx, cost, best = native.topk(_cuda(costs), k, segments)
    idx, cost, best = idx.cpu().numpy(), cost.cpu().numpy(), best.cpu().numpy()
    for s in range(segments):
        want = po.topk_stable(costs[s], k)
        np.testing.assert_array_equal(idx[s], np.sort(want))  # same set, ascending index order
        np.testing.assert_array_equal(cost[s], costs[s][idx[s]])
        assert best[s, 2] == np.argmin(costs[s])  # first minimum, like np.argmin (planners.py:184)
        assert best[s, 0].view(np.float32) == costs[s].min()


@pytest.mark.parametrize("n", [31, 33, 16383, 16384, 16385, 16400, 32768, 32801, 49151, 49152, 49153, 65537])
def test_topk_size_boundaries(native, n):
    """Sizes around the kernel's internal boundaries: the 32-key padding / swizzle groups, the 16384-key
    compaction pass (whose output buffer reuses the pass's own key slice), the 49152-key staging limit
    (beyond it the costs are re-read from global memory), with k = 1, ~10 %, n - 1 and n, on continuous
    costs (two or three radix rounds), heavily tied costs, all-equal costs and all-NaN costs."""
    rng = np.random.default_rng(n)
    cases = {"continuous": (300.0 + 40.0 * rng.standard_normal(n)).astype(np.float32),
             "wide": np.exp(rng.uniform(-30, 30, n)).astype(np.float32) * rng.choice([-1.0, 1.0], n).astype(np.float32),
             "tied": np.round(rng.normal(size=n) * 3).astype(np.float32),
             "equal": np.full(n, 7.25, np.float32),
             "nan": np.full(n, np.nan, np.float32)}
    for name, c in cases.items():
        for k in sorted({1, max(1, n // 10), max(1, n - 1), n}):
            idx, cost, best = native.topk(_cuda(c[None]), k, 1)
            want = np.sort(po.topk_stable(c, k))
            np.testing.assert_array_equal(idx.cpu().numpy()[0], want, err_msg=f"{name} n={n} k={k}")
            np.testing.assert_array_equal(cost.cpu().numpy()[0], c[want], err_msg=f"{name} n={n} k={k}")
            assert best.cpu().numpy()[0, 2] == (0 if name in ("equal", "nan") else np.argmin(c)), f"{name} n={n} k={k}"


def test_topk_special_values(native):
    c = np.array([[3.0, -0.0, 0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 2.0, np.nan, -np.inf, 0.0]], np.float32)
    for k in range(1, c.shape[1] + 1):
        idx, _, best = native.topk(_cuda(c), k, 1)
        np.testing.assert_array_equal(idx.cpu().numpy()[0], np.sort(po.topk_stable(c[0], k)))
    assert best.cpu().numpy()[0, 2] == 4


def test_topk_costs_from_reference(native):
    """Elite indices on the reference's own cost arrays (north_star: bit-exact)."""
    for name in ("cem_cheetah_small.npz", "cem_cartpole_small.npz"):
        g = load_golden(name)
        k = int(g["k"])
        for it in range(int(g["iters"])):
            idx, _, best = native.topk(_cuda(g[f"costs_{it}"]), k, 1)
            np.testing.assert_array_equal(idx.cpu().numpy()[0], np.sort(g[f"elite_{it}"]))
            assert best.cpu().numpy()[0, 2] == g[f"elite_{it}"][0]


# ---------------------------------------------------------------------------------------
# refit + CEM
# ---------------------------------------------------------------------------------------
def test_refit_matches_oracle(native):
    p = po.synthetic_params(6, 5, 16)
    H, n, E, k = 4, 500, 2, 50
    h = _planner(native, p, H, n, E)
    g = torch.Generator().manual_seed(3)
    acts = torch.rand(H * E * n, 5, generator=g) * 2 - 1
    elite = torch.stack([torch.randperm(n, generator=g)[:k].sort().values for _ in range(E)]).int()
    mu, sd = h.refit(elite.cuda(), k, native.SAMPLE_INJECT_ACTIONS, d_injected=acts.cuda())
    a4 = acts.view(H, E, n, 5)
    for e in range(E):
        m, s = po.refit(a4[:, e], elite[e].numpy())
        np.testing.assert_allclose(mu[e].cpu().numpy(), m.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(sd[e].cpu().numpy(), s.numpy(), rtol=1e-5, atol=1e-6)
    # regenerated-from-Philox path == materialised sampler path
    mu0 = (torch.rand(E, H, 5, generator=g) - 0.5).cuda()
    sd0 = (torch.rand(E, H, 5, generator=g) + 0.2).cuda()
    mat = h.sample(native.SAMPLE_GAUSSIAN, 11, 2, mu0, sd0)
    m1, s1 = h.refit(elite.cuda(), k, native.SAMPLE_GAUSSIAN, 11, 2, d_mu=mu0, d_sd=sd0)
    m2, s2 = h.refit(elite.cuda(), k, native.SAMPLE_INJECT_ACTIONS, d_injected=mat)
    # same actions bit for bit; the two paths centre their sums differently (old mean vs 0)
    torch.testing.assert_close(m1, m2, rtol=0, atol=2e-6)
    torch.testing.assert_close(s1, s2, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("name", ["cem_cheetah_small.npz", "cem_cartpole_small.npz"])
def test_cem_plan_matches_reference_composition(native, name):
    g = load_golden(name)
    p = params_from_golden(g)
    n, H, k, I = int(g["n"]), int(g["horizon"]), int(g["k"]), int(g["iters"])
    h = _planner(native, p, H, n, 1, I)
    out = h.plan(g["s0"], I, k, native.SAMPLE_INJECT_NOISE, injected=g["noise"], want_dist=True)
    info = out["info"][0]
    assert (int(info["best_iteration"]), int(info["best_index"])) == (int(g["best_it"]), int(g["best_idx"]))
    np.testing.assert_allclose(info["best_cost"], g["best_cost"], rtol=FP32_COST_RTOL)
    np.testing.assert_allclose(out["actions"][0], g["best_actions"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out["states"][0], g["best_states"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(out["mu"][0], g[f"mu_{I - 1}"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(out["sd"][0], g[f"sd_{I - 1}"], rtol=1e-3, atol=2e-5)


# ---------------------------------------------------------------------------------------
# reference-facing Python API
# ---------------------------------------------------------------------------------------
class _Lin:
    def __init__(self, W, b):
        self.weight, self.bias = torch.nn.Parameter(W.clone()), torch.nn.Parameter(b.clone())


class _Net(torch.nn.Module):  # stands in for src/mbrl/models.py:96 Model
    def __init__(self, p):
        super().__init__()
        self.linear1, self.linear2, self.linear3 = (torch.nn.Linear(1, 1) for _ in range(3))
        for lin, (W, b) in zip((self.linear1, self.linear2, self.linear3), ((p.W1, p.b1), (p.W2, p.b2), (p.W3, p.b3))):
            lin.weight, lin.bias = torch.nn.Parameter(W.clone()), torch.nn.Parameter(b.clone())
        self.noise = None


def _norm(field_value, field_name, stats):
    return (field_value - stats[field_name]["mean"]) / stats[field_name]["std"]


def _unnorm(field_value, field_name, stats):
    return field_value * stats[field_name]["std"] + stats[field_name]["mean"]


class _StateCost:
    def __init__(self, p):
        self.weights, self.goal_state, self.alpha = p.cost_w, p.goal, p.alpha


class _ActCost:
    def __init__(self, p):
        self.alpha = p.beta


class _Spec:
    def __init__(self, a):
        self.minimum, self.maximum, self.shape = np.full(a, -1.0), np.full(a, 1.0), (a,)


def _reference_style_callables(p, recorded):
    stats = {"observations": {"mean": p.mu_s, "std": p.sd_s}, "actions": {"mean": p.mu_a, "std": p.sd_a}}
    model = functools.partial(
        _Net(p),
        normalize_state=functools.partial(_norm, field_name="observations", stats=stats),
        normalize_action=functools.partial(_norm, field_name="actions", stats=stats),
        unnormalize_state=functools.partial(_unnorm, field_name="observations", stats=stats))
    cost = functools.partial(lambda s, a, state_cost, action_cost: None, state_cost=_StateCost(p), action_cost=_ActCost(p))

    def _sample(action_spec, batch_size=None):
        return recorded.clone()

    return model, cost, functools.partial(_sample, action_spec=_Spec(p.act_dim)), stats


def test_python_planner_api_drop_in(native):
    import mbrl_b200
    g = load_golden("rs_cartpole.npz")
    p = params_from_golden(g)
    n, H = int(g["n"]), int(g["horizon"])
    model, cost, sampler, stats = _reference_style_callables(p, torch.from_numpy(g["actions"]))
    s, a = mbrl_b200.RandomShootingPlanner.plan(torch.from_numpy(g["s0"]), model, cost, sampler, H, None,
                                                num_trajectories=n, sampler="host")
    assert s.shape == (H, p.obs_dim) and a.shape == (H, p.act_dim)
    np.testing.assert_array_equal(a.numpy(), g["plan_actions"])
    np.testing.assert_array_equal(a[0].flatten().numpy(), g["first_action"])  # what MPCPolicy returns (agents.py:56)
    np.testing.assert_allclose(s.numpy(), g["plan_states"], rtol=FP32_STATE_RTOL, atol=FP32_STATE_ATOL)
    # host retrains in place -> weights re-uploaded on the next call
    with torch.no_grad():
        model.func.linear3.bias.add_(0.25)
    p2 = params_from_golden(g)
    p2.b3 = p2.b3 + 0.25
    want = po.rs_plan(p2, torch.from_numpy(g["s0"]), torch.from_numpy(g["actions"]), H, n)
    s2, a2 = mbrl_b200.RandomShootingPlanner.plan(torch.from_numpy(g["s0"]), model, cost, sampler, H, None,
                                                  num_trajectories=n, sampler="host")
    np.testing.assert_array_equal(a2.numpy(), want["actions"].numpy())
    np.testing.assert_allclose(s2.numpy(), want["states"].numpy(), rtol=FP32_STATE_RTOL, atol=FP32_STATE_ATOL)
    # device sampler + CEM run and respect the action bounds
    s3, a3 = mbrl_b200.CEMPlanner.plan(torch.from_numpy(g["s0"]), model, cost, sampler, H, None,
                                       num_trajectories=2048, num_iterations=3)
    assert a3.abs().max() <= 1.0 and torch.isfinite(s3).all()
    with pytest.raises(TypeError):
        mbrl_b200.RandomShootingPlanner.plan(torch.from_numpy(g["s0"]), lambda s, a: s, cost, sampler, H)
    mbrl_b200.planners.clear_handles()


# ---------------------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties
# ---------------------------------------------------------------------------------------
# per-engine tolerances of trajectory costs against the fp32 oracle (tests/test_gpu_tc.py states them)
ENGINE_COST_RTOL = {"fp32": FP32_COST_RTOL, "fp16": 1e-4, "bf16": 5e-4}


@pytest.mark.parametrize("engine", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("shape", [(17, 6, 200), (24, 6, 200)], ids=["cheetah_cfg3", "walker_cfg4_shard"])
def test_full_size_cem_properties(native, engine, shape):
    """BASELINE cfg 3 (cheetah-run shape) and one cfg-4 shard (walker-walk shape), N=16384 H=30 I=5
    k=1638, on the device sampler, on EVERY engine (the benchmarked fp16 engine included): rollout costs
    against the oracle on a candidate subset with the device's own draws injected, bit-exact elite
    selection on the engine's own cost array, refit against the oracle, monotone best-ever cost, and a
    plan whose reported cost the oracle reproduces."""
    O, A, U = shape
    H, N, I, k = 30, 16384, 5, 1638
    rtol = ENGINE_COST_RTOL[engine]
    p = po.synthetic_params(O, A, U)
    s0 = po.synthetic_state(p, 0)
    h = _planner(native, p, H, N, 1, I, engine=engine)
    mu = torch.zeros(1, H, A, device="cuda")
    sd = torch.ones(1, H, A, device="cuda")
    d_s0 = s0[None].cuda()
    best_prev = np.inf
    for it in range(I):
        costs, _, acts = h.rollout(d_s0, native.SAMPLE_GAUSSIAN, 42, it, d_mu=mu, d_sd=sd, want_actions=True)
        # oracle on 96 candidates spread over the population, with the device's own draws injected
        sub = torch.arange(0, N, N // 96)[:96]
        a_sub = acts.cpu().view(H, N, A)[:, sub].reshape(H * 96, A)
        _, c_ref = po.rollout_costs(p, s0, a_sub, H, 96)
        np.testing.assert_allclose(costs.cpu().numpy()[sub.numpy()], c_ref, rtol=rtol)
        idx, ecost, best = native.topk(costs, k, 1)
        e = idx.cpu().numpy()[0]
        assert len(np.unique(e)) == k and (np.diff(e) > 0).all()
        c = costs.cpu().numpy()
        np.testing.assert_array_equal(e, np.sort(po.topk_stable(c, k)))
        assert c[e].max() <= np.delete(c, e).min()
        mu, sd = h.refit(idx, k, native.SAMPLE_GAUSSIAN, 42, it, d_mu=mu, d_sd=sd)
        m_ref, s_ref = po.refit(acts.cpu().view(H, N, A), e)
        np.testing.assert_allclose(mu[0].cpu().numpy(), m_ref.numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(sd[0].cpu().numpy(), s_ref.numpy(), rtol=1e-4, atol=1e-5)
        best_prev = min(best_prev, float(c.min()))
    out = h.plan(s0.numpy(), I, k, native.SAMPLE_GAUSSIAN, 42)
    np.testing.assert_allclose(out["info"]["best_cost"][0], best_prev, rtol=1e-6)
    # replayed plan reproduces its own cost under the oracle
    _, c_plan = po.rollout_costs(p, s0, torch.from_numpy(out["actions"][0]), H, 1)
    np.testing.assert_allclose(c_plan[0], out["info"]["best_cost"][0], rtol=rtol)


@pytest.mark.parametrize("engine", ["fp16", "bf16"])
@pytest.mark.parametrize("name", ["cem_cheetah_small.npz", "cem_cartpole_small.npz"])
def test_tensor_core_cem_with_reference_noise(native, engine, name):
    """north_star's third parity clause on the benchmarked engine: CEM with the REFERENCE-injected noise
    (the fixture recorded the draws the reference composition consumed).  Per iteration, with the
    reference's own sampling distribution: costs within the engine's tolerance and the elite set against
    the reference's (overlap; on identical cost arrays the select is bit-exact, test_topk_costs_from_reference).
    Whole plan: best cost within tolerance, the same winning candidate, first planned action within
    FIRST_ACTION_ATOL of the reference's (actions live in [-1, 1])."""
    FIRST_ACTION_ATOL = 0.05
    MIN_OVERLAP = {"fp16": 0.88, "bf16": 0.80}[engine]
    g = load_golden(name)
    p = params_from_golden(g)
    n, H, k, I, A = int(g["n"]), int(g["horizon"]), int(g["k"]), int(g["iters"]), p.act_dim
    h = _planner(native, p, H, n, 1, I, engine=engine)
    d_s0 = _cuda(g["s0"][None])
    lo, hi = float(g["lo"]), float(g["hi"])
    report = []
    for it in range(I):
        mu = np.full((1, H, A), 0.5 * (lo + hi), np.float32) if it == 0 else g[f"mu_{it - 1}"][None]
        sd = np.full((1, H, A), 0.5 * (hi - lo), np.float32) if it == 0 else g[f"sd_{it - 1}"][None]
        costs, _, _ = h.rollout(d_s0, native.SAMPLE_INJECT_NOISE, d_injected=_cuda(g["noise"][it]), d_mu=_cuda(mu), d_sd=_cuda(sd))
        c = costs.cpu().numpy()
        rel = np.abs(c - g[f"costs_{it}"]) / np.abs(g[f"costs_{it}"])
        idx, _, _ = native.topk(costs, k, 1)
        overlap = len(set(idx.cpu().numpy()[0].tolist()) & set(g[f"elite_{it}"].tolist())) / k
        report.append(f"{engine} {name} it {it}: cost rel err {rel.max():.2e}, elite overlap {overlap:.2f}")
        assert rel.max() <= ENGINE_COST_RTOL[engine], report[-1]
        assert overlap >= MIN_OVERLAP, report[-1]
    out = h.plan(g["s0"], I, k, native.SAMPLE_INJECT_NOISE, injected=g["noise"], want_dist=True)
    info = out["info"][0]
    a0_err = np.abs(out["actions"][0][0] - g["best_actions"][0]).max()
    cost_err = abs(float(info["best_cost"]) - float(g["best_cost"])) / abs(float(g["best_cost"]))
    report.append(f"{engine} {name} plan: best (it, idx) = ({int(info['best_iteration'])}, {int(info['best_index'])}) vs reference "
                  f"({int(g['best_it'])}, {int(g['best_idx'])}); best cost rel err {cost_err:.2e}; first action abs err {a0_err:.2e}")
    print("\n".join(report))
    assert cost_err <= 10 * ENGINE_COST_RTOL[engine], report[-1]  # (the winner may be a near-tie neighbour)
    assert a0_err <= FIRST_ACTION_ATOL, report[-1]


def test_warm_start_and_return_mean(native):
    """Warm start (mu0/sd0, what CEMPlanner builds from MPCPolicy's initial_trajectory,
    src/mbrl/agents.py:41-47) and return_mean: checked against the oracle CEM with the device's own
    draws injected."""
    p = po.synthetic_params(9, 3, 40, seed=6)
    H, n, I, k = 8, 1024, 3, 100
    s0 = po.synthetic_state(p, 1)
    g = torch.Generator().manual_seed(8)
    mu0 = torch.rand(H, 3, generator=g) * 0.6 - 0.3
    sd0 = torch.rand(H, 3, generator=g) * 0.3 + 0.1
    h = _planner(native, p, H, n, 1, I)
    out = h.plan(s0.numpy(), I, k, native.SAMPLE_GAUSSIAN, seed=4, mu0=mu0[None].numpy(), sd0=sd0[None].numpy(), want_dist=True)
    # recover the N(0,1) draws of every iteration from the materialising sampler (mu=0, sd=1, wide bounds)
    h.set_action_bounds(-1e6, 1e6)
    z = torch.stack([h.sample(native.SAMPLE_GAUSSIAN, 4, it, torch.zeros(1, H, 3, device="cuda"),
                              torch.ones(1, H, 3, device="cuda")).cpu() for it in range(I)])
    h.set_action_bounds(p.act_lo, p.act_hi)
    ref = po.cem_plan(p, s0, z, H, n, k, mu0=mu0, sd0=sd0)
    assert (int(out["info"]["best_iteration"][0]), int(out["info"]["best_index"][0])) == (ref["best"]["it"], ref["best"]["idx"])
    np.testing.assert_allclose(out["actions"][0], ref["best"]["actions"].numpy(), atol=2e-6)
    np.testing.assert_allclose(out["mu"][0], ref["mu"].numpy(), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(out["sd"][0], ref["sd"].numpy(), rtol=1e-3, atol=2e-5)
    # return_mean: the emitted sequence is the final mean, replayed through the model
    outm = h.plan(s0.numpy(), I, k, native.SAMPLE_GAUSSIAN, seed=4, mu0=mu0[None].numpy(), sd0=sd0[None].numpy(),
                  return_mean=True, want_dist=True)
    np.testing.assert_allclose(outm["actions"][0], outm["mu"][0], atol=1e-7)
    st, _ = po.rollout_costs(p, s0, torch.from_numpy(outm["actions"][0]), H, 1)
    np.testing.assert_allclose(outm["states"][0], st.numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("engine", ["fp32", "fp16"])
def test_device_resident_warm_start_equals_host_round_trip(native, engine):
    """MBRL_WARM_KEEP / MBRL_WARM_USE (the mean stays on the device between MPC steps,
    src/mbrl/agents.py:41-47): plan 2 warm-started on the device equals, bit for bit, plan 2 seeded from the
    host with plan 1's final mean shifted by one step; and a call without MBRL_WARM_KEEP forgets the mean."""
    p = po.synthetic_params(17, 6, 200)
    H, N, I, k = 12, 1024, 3, 102
    s0a, s0b = po.synthetic_state(p, 1).numpy(), po.synthetic_state(p, 2).numpy()
    h = _planner(native, p, H, N, 1, I, engine=engine)
    first = h.plan(s0a, I, k, native.SAMPLE_GAUSSIAN, seed=1, want_dist=True, warm_start=native.WARM_KEEP)
    dev = h.plan(s0b, I, k, native.SAMPLE_GAUSSIAN, seed=2, want_dist=True, warm_start=native.WARM_KEEP | native.WARM_USE, warm_std=0.4)
    mu0 = np.concatenate([first["mu"][0][1:], first["mu"][0][-1:]])[None]
    h2 = _planner(native, p, H, N, 1, I, engine=engine)
    host = h2.plan(s0b, I, k, native.SAMPLE_GAUSSIAN, seed=2, want_dist=True, mu0=mu0, sd0=np.full_like(mu0, 0.4))
    for key in ("actions", "states", "mu", "sd"):
        np.testing.assert_array_equal(dev[key], host[key], err_msg=key)
    assert dev["info"]["best_cost"][0] == host["info"]["best_cost"][0]
    # a plan without KEEP drops the stored mean: the next USE call is a cold start
    h.plan(s0a, I, k, native.SAMPLE_GAUSSIAN, seed=3)
    cold = h.plan(s0b, I, k, native.SAMPLE_GAUSSIAN, seed=2, want_dist=True, warm_start=native.WARM_USE)
    ref = h2.plan(s0b, I, k, native.SAMPLE_GAUSSIAN, seed=2, want_dist=True)
    np.testing.assert_array_equal(cold["actions"], ref["actions"])


def test_refit_large_elite_set_chunked(native):
    """k > 2048 elites: the refit runs as parallel 2048-elite chunks whose partial sums are added in
    chunk order -- checked against the oracle refit and for run-to-run bit stability."""
    p = po.synthetic_params(6, 5, 16)
    H, n, E, k = 3, 9000, 2, 5000
    h = _planner(native, p, H, n, E, 1, max_elites=k)
    g = torch.Generator().manual_seed(5)
    acts = torch.rand(H * E * n, 5, generator=g) * 2 - 1
    elite = torch.stack([torch.randperm(n, generator=g)[:k].sort().values for _ in range(E)]).int()
    mu, sd = h.refit(elite.cuda(), k, native.SAMPLE_INJECT_ACTIONS, d_injected=acts.cuda())
    a4 = acts.view(H, E, n, 5)
    for e in range(E):
        m, s = po.refit(a4[:, e], elite[e].numpy())
        np.testing.assert_allclose(mu[e].cpu().numpy(), m.numpy(), rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(sd[e].cpu().numpy(), s.numpy(), rtol=1e-5, atol=2e-6)
    for _ in range(5):  # chunk CTAs finish in any order; the sum order must not depend on it
        mu2, sd2 = h.refit(elite.cuda(), k, native.SAMPLE_INJECT_ACTIONS, d_injected=acts.cuda())
        assert torch.equal(mu, mu2) and torch.equal(sd, sd2)
    mu0 = (torch.rand(E, H, 5, generator=g) - 0.5).cuda()
    sd0 = (torch.rand(E, H, 5, generator=g) + 0.2).cuda()
    mat = h.sample(native.SAMPLE_GAUSSIAN, 11, 2, mu0, sd0)
    m1, s1 = h.refit(elite.cuda(), k, native.SAMPLE_GAUSSIAN, 11, 2, d_mu=mu0, d_sd=sd0)
    m2, s2 = h.refit(elite.cuda(), k, native.SAMPLE_INJECT_ACTIONS, d_injected=mat)
    torch.testing.assert_close(m1, m2, rtol=0, atol=2e-6)
    torch.testing.assert_close(s1, s2, rtol=1e-5, atol=2e-6)


def test_reward_head_cost_matches_reference_fixture(native):
    """RewardAgent's cost (ModelWithReward's reward head, a second trunk evaluation per step) on the
    fp32 engine and on the tensor-core engines against the fixture produced by the reference's own planner
    (tests/golden/make_golden.py::make_rs_reward): costs, argmin, the returned plan; then the Python
    drop-in API with callables wired exactly like src/mbrl/agents.py:349-358."""
    from functools import partial
    from operator import itemgetter
    from mbrl_b200 import RandomShootingPlanner, planners
    g = load_golden("rs_reward_head.npz")
    p = params_from_golden(g)
    n, H = int(g["n"]), int(g["horizon"])
    h = native.NativePlanner(p.obs_dim, p.act_dim, p.hidden, H, n, 1, 1, None, "fp32")
    h.set_weights(p.W1, p.b1, p.W2, p.b2, p.W3, p.b3)
    h.set_norm(p.mu_s, p.sd_s, p.mu_a, p.sd_a)
    h.set_reward_head(p.W4, p.b4, p.mu_r, p.sd_r)
    h.set_cost(kind=native.COST_REWARD_HEAD)
    h.set_action_bounds(p.act_lo, p.act_hi)
    costs, _, _ = h.rollout(_cuda(g["s0"][None]), native.SAMPLE_INJECT_ACTIONS, d_injected=_cuda(g["actions"]))
    np.testing.assert_allclose(costs.cpu().numpy(), g["costs"], rtol=1e-5, atol=1e-4)
    out = h.plan(g["s0"], 1, 1, native.SAMPLE_INJECT_ACTIONS, injected=g["actions"])
    assert int(out["info"]["best_index"][0]) == int(g["idx"])
    np.testing.assert_array_equal(out["actions"][0], g["plan_actions"])
    np.testing.assert_allclose(out["states"][0], g["plan_states"], rtol=1e-5, atol=1e-5)
    # tensor-core engines: the handle switches to the weight-streaming kernel, which runs the second trunk
    # pass on the tensor cores (16-bit operands: costs within the engine's tolerance of the reference's;
    # the candidate it picks is a minimiser of the reference's cost array up to that tolerance)
    for engine, tol in (("fp16", 2e-3), ("bf16", 1e-2)):
        tc = native.NativePlanner(p.obs_dim, p.act_dim, p.hidden, H, n, 1, 1, None, engine)
        tc.set_reward_head(p.W4, p.b4, p.mu_r, p.sd_r)
        tc.set_weights(p.W1, p.b1, p.W2, p.b2, p.W3, p.b3)   # weights before the cost kind: the image is re-packed on the switch
        tc.set_norm(p.mu_s, p.sd_s, p.mu_a, p.sd_a)
        tc.set_cost(kind=native.COST_REWARD_HEAD)
        tc.set_action_bounds(p.act_lo, p.act_hi)
        c16, _, _ = tc.rollout(_cuda(g["s0"][None]), native.SAMPLE_INJECT_ACTIONS, d_injected=_cuda(g["actions"]))
        err = np.abs(c16.cpu().numpy() - g["costs"]).max()
        scale = np.abs(g["costs"]).max()
        print(f"reward head on {engine}: max abs cost err {err:.3e} (scale {scale:.3f})")
        assert err <= tol * scale + tol, (engine, err)
        o16 = tc.plan(g["s0"], 1, 1, native.SAMPLE_INJECT_ACTIONS, injected=g["actions"])
        assert g["costs"][int(o16["info"]["best_index"][0])] <= g["costs"].min() + 2 * (tol * scale + tol)
        tc.close()

    # drop-in API with the RewardAgent wiring
    class ModelWithReward(torch.nn.Module):
        def __init__(self):
            super().__init__()
            U, O, A = p.hidden, p.obs_dim, p.act_dim
            self.linear1, self.linear2 = torch.nn.Linear(O + A, U), torch.nn.Linear(U, U)
            self.linear3, self.linear4 = torch.nn.Linear(U, O), torch.nn.Linear(U, 1)
    net = ModelWithReward()
    with torch.no_grad():
        for lin, (w, b) in zip((net.linear1, net.linear2, net.linear3, net.linear4),
                               ((p.W1, p.b1), (p.W2, p.b2), (p.W3, p.b3), (p.W4, p.b4))):
            lin.weight.copy_(w); lin.bias.copy_(b)
    stats = {"observations": {"mean": p.mu_s, "std": p.sd_s}, "actions": {"mean": p.mu_a, "std": p.sd_a},
             "rewards": {"mean": torch.tensor([p.mu_r]), "std": torch.tensor([p.sd_r])}}
    def field(x, field_name, stats):  # stands in for TransitionsDataset.(un)normalize_field
        raise AssertionError("the GPU planner never calls the host callables")
    def compose(a, b):
        def ab(*args, **kwargs):
            return b(a(*args, **kwargs))
        return ab
    wired = partial(net, normalize_state=partial(field, field_name="observations", stats=stats),
                    normalize_action=partial(field, field_name="actions", stats=stats),
                    unnormalize_state=partial(field, field_name="observations", stats=stats),
                    unnormalize_reward=partial(field, field_name="rewards", stats=stats))
    acts = torch.from_numpy(g["actions"])
    s, a = RandomShootingPlanner.plan(torch.from_numpy(g["s0"]), compose(wired, itemgetter(0)), compose(wired, itemgetter(1)),
                                      lambda batch_size: acts, H, None, num_trajectories=n, sampler="host", engine="fp32")
    np.testing.assert_array_equal(a.numpy(), g["plan_actions"])
    np.testing.assert_allclose(s.numpy(), g["plan_states"], rtol=1e-5, atol=1e-5)
    planners.clear_handles()


def test_dmc_humanoid_task_cost_epilogue(native):
    """Humanoid-run task cost (dm_control/suite/humanoid.py:187-211 restated on
    the 67-d egocentric observation and the 21-d control) at the cfg-5 model shape on the fp32
    engine, against the oracle that is pinned to the reference's rewards.tolerance composition."""
    from oracle import task_costs
    p = po.synthetic_params(67, 21, 128, seed=4)
    # statistics that put head height / uprightness / speed in the interesting range of the reward
    p.mu_s[21], p.sd_s[21] = 1.2, 0.4
    p.mu_s[36], p.sd_s[36] = 0.5, 0.6
    p.mu_s[37], p.sd_s[37] = 4.0, 5.0
    p.mu_s[38], p.sd_s[38] = 0.0, 3.0
    H, n = 6, 300
    h = _planner(native, p, H, n)
    h.set_cost(kind=native.COST_DMC_HUMANOID_RUN)
    g = torch.Generator().manual_seed(2)
    s0 = po.synthetic_state(p, 3)
    acts = torch.rand(H * n, 21, generator=g) * 2.6 - 1.3  # beyond +-1: the quadratic control tolerance saturates
    h.set_action_bounds(-1.3, 1.3)
    costs, states, _ = h.rollout(s0[None].cuda(), native.SAMPLE_INJECT_ACTIONS, d_injected=acts.cuda(), want_states=True)
    st, _ = po.rollout_costs(p, s0, acts, H, n)
    np.testing.assert_allclose(states.cpu().numpy(), st.numpy(), rtol=2e-4, atol=2e-4)
    want = task_costs.humanoid_cost(st.numpy(), acts.numpy()).reshape(H, n).sum(0)
    assert want.std() > 0.01  # the test exercises the reward's active range
    np.testing.assert_allclose(costs.cpu().numpy(), want, rtol=1e-4, atol=1e-4)
    with pytest.raises(native.MbrlError):
        _planner(native, po.synthetic_params(17, 6, 64), H, n).set_cost(kind=native.COST_DMC_HUMANOID_RUN)  # obs too small


def test_linear_model_drop_in_matches_reference_fixture(native):
    """`--model lin` (LinearModel, src/mbrl/models.py:113-122) through the drop-in planner API: the
    adaptor's exact ReLU embedding on the fp32 engine reproduces the reference planner's plan."""
    from functools import partial
    from mbrl_b200 import RandomShootingPlanner, planners
    g = load_golden("rs_linear_model.npz")
    n, H = int(g["n"]), int(g["horizon"])
    O, D = g["W1"].shape

    class LinearModel(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.linear1 = torch.nn.Linear(D, O)
            self.noise = None
    net = LinearModel()
    with torch.no_grad():
        net.linear1.weight.copy_(torch.from_numpy(g["W1"])); net.linear1.bias.copy_(torch.from_numpy(g["b1"]))
    t = lambda k: torch.from_numpy(g[k])
    stats = {"observations": {"mean": t("mu_s"), "std": t("sd_s")}, "actions": {"mean": t("mu_a"), "std": t("sd_a")}}

    def field(x, field_name, stats):
        raise AssertionError("the GPU planner never calls the host callables")

    class SC:
        weights, goal_state, alpha = t("cost_w"), t("goal"), float(g["alpha"])

    class AC:
        alpha = float(g["beta"])
    model = partial(net, normalize_state=partial(field, field_name="observations", stats=stats),
                    normalize_action=partial(field, field_name="actions", stats=stats),
                    unnormalize_state=partial(field, field_name="observations", stats=stats))
    cost = partial(field, state_cost=SC, action_cost=AC)
    acts = torch.from_numpy(g["actions"])
    s, a = RandomShootingPlanner.plan(t("s0"), model, cost, lambda batch_size: acts, H, None, num_trajectories=n,
                                      sampler="host", engine="fp32")
    np.testing.assert_array_equal(a.numpy(), g["plan_actions"])
    np.testing.assert_allclose(s.numpy(), g["plan_states"], rtol=1e-5, atol=1e-5)
    s16, a16 = RandomShootingPlanner.plan(t("s0"), model, cost, lambda batch_size: acts, H, None, num_trajectories=n,
                                          sampler="host", engine="fp16", return_states=True)
    # fp16 engine: the candidate it picks must be a minimiser of the REFERENCE's cost array within the
    # engine's cost tolerance (a near-tie may legitimately resolve differently), and the states it
    # returns are the fp32 replay of that candidate
    assert a16.shape == a.shape and torch.isfinite(s16).all()
    cand = acts.view(H, n, -1).permute(1, 0, 2)  # [n, H, A]
    match = (cand == a16[None]).all(dim=2).all(dim=1).nonzero().flatten()
    assert match.numel() >= 1, "the fp16 plan is not one of the injected candidates"
    ref_costs = g["costs"]
    assert ref_costs[int(match[0])] <= ref_costs.min() * (1 + 1e-3) + 1e-6
    if int(match[0]) == int(g["idx"]):
        np.testing.assert_allclose(s16.numpy(), g["plan_states"], rtol=1e-5, atol=1e-5)
    planners.clear_handles()


# ---------------------------------------------------------------------------------------
# GradientDescentPlanner
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["gd_small.npz", "gd_early_stop.npz", "gd_cheetah.npz"])
def test_gradient_descent_planner_matches_reference_fixture(native, name):
    """The batched GPU GradientDescentPlanner against the reference's own planner run as shipped
    (src/mbrl/planners.py:28-137; tests/golden/make_golden.py::make_gd): same iteration count (incl. the early
    stop), returned actions and states.  Tolerance: Adam normalises the gradient, so a step is ~lr * sign(g)
    whatever the gradient's magnitude; fp32 summation-order differences (analytic back-propagation vs torch
    autograd + MKL) stay far below lr = 0.01 -- actions within 2e-4, states within 1e-3."""
    g = load_golden(name)
    p = params_from_golden(g)
    H = int(g["horizon"])
    h = _planner(native, p, H, 1, 1, 1, engine="fp32")
    out = h.plan_gd(g["s0"], g["init_actions"][None], iterations=int(g["iters"]), stop_condition=float(g["stop"]))
    assert int(out["iterations"][0]) == int(g["iterations_run"])
    err_a = np.abs(out["actions"][0] - g["actions"]).max()
    err_s = np.abs(out["states"][0] - g["states"]).max()
    print(f"GD planner {name}: iterations {int(out['iterations'][0])}, max |action err| {err_a:.2e}, max |state err| {err_s:.2e}")
    assert err_a <= 2e-4 and err_s <= 1e-3
    np.testing.assert_array_equal(out["states"][0][0], g["s0"])


def test_gradient_descent_planner_batched_restarts_and_api(native):
    """Batched restarts: every restart is optimised independently (restart 0 equals the single-restart run bit for
    bit), the loss goes down, and the Python drop-in class returns the reference's list-of-[1, .]-tensors shape
    (planners.py:137) with the best restart."""
    from functools import partial
    from mbrl_b200 import GradientDescentPlanner, planners
    g = load_golden("gd_small.npz")
    p = params_from_golden(g)
    H, A, O = int(g["horizon"]), p.act_dim, p.obs_dim
    h = _planner(native, p, H, 1, 1, 1, engine="fp32")
    rng = np.random.default_rng(0)
    init = np.concatenate([g["init_actions"][None], rng.uniform(-1, 1, (6, H, A)).astype(np.float32)])
    one = h.plan_gd(g["s0"], init[:1], iterations=25)
    many = h.plan_gd(g["s0"], init, iterations=25)
    np.testing.assert_array_equal(many["actions"][0], one["actions"][0])
    for b in range(init.shape[0]):
        _, c0 = po.rollout_costs(p, torch.from_numpy(g["s0"]), torch.from_numpy(init[b]), H, 1)
        assert many["cost"][b] < c0[0], "25 Adam iterations must lower the trajectory cost"
    # drop-in class, callables wired like GoalStateAgent (agents.py:225-233)
    model, cost, _, _ = _reference_style_callables(p, None)
    calls = []

    def sample(batch_size):
        calls.append(batch_size)
        return torch.from_numpy(init[len(calls) - 1])
    states, actions = GradientDescentPlanner.plan(torch.from_numpy(g["s0"]), model, cost, sample, H, None, num_iterations=25,
                                                  num_restarts=4)
    assert calls == [H] * 4 and len(states) == H + 1 and len(actions) == H
    assert states[0].shape == (1, O) and actions[0].shape == (1, A)
    best = int(np.argmin(many["cost"][:4]))
    np.testing.assert_allclose(torch.cat(actions).numpy(), many["actions"][best], rtol=0, atol=1e-6)
    assert actions[0].flatten().shape == (A,)  # what MPCPolicy.get_action returns (agents.py:56)
    planners.clear_handles()
