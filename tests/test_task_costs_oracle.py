"""Pins oracle/task_costs.tolerance to values produced by the reference's own
dm_control/utils/rewards.py (tests/golden/tolerance.npz, made by tests/golden/make_golden.py) and to
the known answers asserted in dm_control/utils/rewards_test.py:39-103."""
import numpy as np
import pytest

from mbrl_helpers import load_golden
from oracle import task_costs as tc

SIGMOIDS = ["gaussian", "linear", "quadratic", "hyperbolic", "long_tail", "cosine", "tanh_squared"]


def test_tolerance_matches_reference_grid():
    g = load_golden("tolerance.npz")
    for case, want in zip(g["cases"], g["values"]):
        sig, lo, hi, margin, vam = SIGMOIDS[int(case[0])], case[1], case[2], case[3], case[4]
        got = tc.tolerance(g["x"], bounds=(lo, hi), margin=margin, sigmoid=sig, value_at_margin=vam)
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-15, err_msg=str(case))


@pytest.mark.parametrize("sigmoid", SIGMOIDS)
def test_tolerance_known_answers(sigmoid):
    """rewards_test.py: value 1 inside the bounds, value_at_margin exactly at the margin."""
    vam = 0.3
    lo, hi, margin = -0.5, 1.5, 2.0
    assert tc.tolerance(np.array([lo, 0.3, hi]), (lo, hi), margin, sigmoid, vam).tolist() == [1.0, 1.0, 1.0]
    at_margin = tc.tolerance(np.array([lo - margin, hi + margin]), (lo, hi), margin, sigmoid, vam)
    np.testing.assert_allclose(at_margin, vam, rtol=1e-9)
    assert tc.tolerance(np.array([-9.0, 9.0]), (lo, hi), 0.0, sigmoid, vam).tolist() == [0.0, 0.0]


def test_cartpole_cost_range_and_optimum():
    upright = np.array([0.0, 1.0, 0.0, 0.0, 0.0])
    assert tc.cartpole_swingup_cost(upright, np.array([0.0])) == pytest.approx(0.0, abs=1e-12)
    hanging = np.array([0.0, -1.0, 0.0, 0.0, 0.0])
    assert tc.cartpole_swingup_cost(hanging, np.array([0.0])) == pytest.approx(1.0, abs=1e-12)
    rng = np.random.default_rng(0)
    c = tc.cartpole_swingup_cost(rng.normal(size=(1000, 5)).clip(-3, 3) * [1, 0.3, 0.3, 1, 3], rng.uniform(-1, 1, (1000, 1)))
    assert ((c >= 0) & (c <= 1.0 + 1e-12)).all()


def test_humanoid_cost_matches_reference_composition():
    """oracle humanoid cost == 1 - Humanoid.get_reward composed with the reference's own
    rewards.tolerance (tests/golden/humanoid_reward.npz, humanoid.py:187-211, move_speed = 10)."""
    g = load_golden("humanoid_reward.npz")
    m = g["reward"].shape[0]
    obs = np.zeros((m, 67))
    obs[:, 21] = g["head_height"]
    obs[:, 36] = g["torso_upright"]
    obs[:, 37:40] = g["com_velocity"]
    np.testing.assert_allclose(tc.humanoid_cost(obs, g["control"]), 1.0 - g["reward"], rtol=1e-12, atol=1e-14)
    assert 0.0 <= (1.0 - g["reward"]).min() and (1.0 - g["reward"]).max() <= 1.0


def test_cheetah_and_walker_costs_match_reference_composition():
    """oracle cheetah-run / walker-walk costs == 1 - the tasks' get_reward composed with the reference's own
    rewards.tolerance (tests/golden/locomotion_reward.npz; cheetah.py:91-97, walker.py:135-158) with the
    speed proxies in the documented observation slots."""
    g = load_golden("locomotion_reward.npz")
    m = g["cheetah_reward"].shape[0]
    obs_c = np.zeros((m, 17))
    obs_c[:, 8] = g["cheetah_speed"]
    np.testing.assert_allclose(tc.cheetah_run_cost(obs_c), 1.0 - g["cheetah_reward"], rtol=1e-12, atol=1e-14)
    obs_w = np.zeros((m, 24))
    obs_w[:, 14], obs_w[:, 0], obs_w[:, 16] = g["walker_height"], g["walker_upright"], g["walker_velocity"]
    np.testing.assert_allclose(tc.walker_walk_cost(obs_w), 1.0 - g["walker_reward"], rtol=1e-12, atol=1e-14)
    for c in (1.0 - g["cheetah_reward"], 1.0 - g["walker_reward"]):
        assert 0.0 <= c.min() and c.max() <= 1.0 and c.std() > 0.05
