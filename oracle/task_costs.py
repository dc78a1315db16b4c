"""dm_control task costs restated on observations (TEST INFRASTRUCTURE).

The reference planner never calls these (the environment computes rewards from MuJoCo physics,
dm_control/dm_control/rl/control.py:111); they are the north_star's optional "task cost"
epilogue.  ``tolerance`` restates dm_control/dm_control/utils/rewards.py:28-130 and is pinned
against values produced by the reference module itself (tests/golden/tolerance.npz)."""
import numpy as np

_DEFAULT_VALUE_AT_MARGIN = 0.1  # rewards.py:25


def _sigmoid(x, value_at_1, kind):
    """rewards.py:28-85: 1 at x == 0, value_at_1 at x == 1."""
    if kind == "gaussian":
        scale = np.sqrt(-2 * np.log(value_at_1))
        return np.exp(-0.5 * (x * scale) ** 2)
    if kind == "hyperbolic":
        scale = np.arccosh(1 / value_at_1)
        return 1 / np.cosh(x * scale)
    if kind == "long_tail":
        scale = np.sqrt(1 / value_at_1 - 1)
        return 1 / ((x * scale) ** 2 + 1)
    if kind == "cosine":
        scale = np.arccos(2 * value_at_1 - 1) / np.pi
        sx = x * scale
        with np.errstate(invalid="ignore"):
            return np.where(np.abs(sx) < 1, (1 + np.cos(np.pi * sx)) / 2, 0.0)
    if kind == "linear":
        sx = x * (1 - value_at_1)
        return np.where(np.abs(sx) < 1, 1 - sx, 0.0)
    if kind == "quadratic":
        sx = x * np.sqrt(1 - value_at_1)
        return np.where(np.abs(sx) < 1, 1 - sx ** 2, 0.0)
    if kind == "tanh_squared":
        scale = np.arctanh(np.sqrt(1 - value_at_1))
        return 1 - np.tanh(x * scale) ** 2
    raise ValueError(f"unknown sigmoid {kind!r}")


def tolerance(x, bounds=(0.0, 0.0), margin=0.0, sigmoid="gaussian", value_at_margin=_DEFAULT_VALUE_AT_MARGIN):
    """rewards.py:88-130: 1 inside [lower, upper], sigmoidal decay over `margin` outside."""
    x = np.asarray(x, dtype=np.float64)
    lower, upper = bounds
    in_bounds = np.logical_and(lower <= x, x <= upper)
    if margin == 0:
        return np.where(in_bounds, 1.0, 0.0)
    d = np.where(x < lower, lower - x, x - upper) / margin
    return np.where(in_bounds, 1.0, _sigmoid(d, value_at_margin, sigmoid))


def cartpole_swingup_cost(obs, act):
    """1 - smooth cartpole reward (dm_control/dm_control/suite/cartpole.py:216-226) from the
    observation [x, cos(theta), sin(theta), x_dot, theta_dot] (cartpole.py:150-153, 202-207) and
    the control.  obs [..., 5], act [..., 1] -> cost [...]."""
    obs = np.asarray(obs, dtype=np.float64)
    act = np.asarray(act, dtype=np.float64)
    upright = (obs[..., 1] + 1) / 2
    centered = (1 + tolerance(obs[..., 0], margin=2)) / 2
    small_control = (4 + tolerance(act[..., 0], margin=1, value_at_margin=0, sigmoid="quadratic")) / 5
    small_velocity = (1 + tolerance(obs[..., 4], margin=5)) / 2
    return 1.0 - upright * small_control * small_velocity * centered


HUMANOID_STAND_HEIGHT = 1.4  # dm_control/suite/humanoid.py:36
HUMANOID_RUN_SPEED = 10.0    # dm_control/suite/humanoid.py:40


def humanoid_cost(obs, act, move_speed=HUMANOID_RUN_SPEED):
    """1 - Humanoid.get_reward (dm_control/dm_control/suite/humanoid.py:187-211) from the
    egocentric observation (humanoid.py:172-185): joint_angles[0:21], head_height[21],
    extremities[22:34], torso_vertical[34:37] (zz = obs[36] is physics.torso_upright(),
    humanoid.py:95-97), com_velocity[37:40], velocity[40:67]; control = the action.
    move_speed > 0 (walk / run; the reference's stand task, move_speed == 0, is `dont_move`).
    obs [..., 67], act [..., 21] -> cost [...]."""
    obs = np.asarray(obs, dtype=np.float64)
    act = np.asarray(act, dtype=np.float64)
    standing = tolerance(obs[..., 21], bounds=(HUMANOID_STAND_HEIGHT, np.inf), margin=HUMANOID_STAND_HEIGHT / 4)
    upright = tolerance(obs[..., 36], bounds=(0.9, np.inf), sigmoid="linear", margin=1.9, value_at_margin=0)
    small_control = tolerance(act, margin=1, value_at_margin=0, sigmoid="quadratic").mean(axis=-1)
    small_control = (4 + small_control) / 5
    com_velocity = np.sqrt(obs[..., 37] ** 2 + obs[..., 38] ** 2)
    move = tolerance(com_velocity, bounds=(move_speed, np.inf), margin=move_speed, value_at_margin=0, sigmoid="linear")
    move = (5 * move + 1) / 6
    return 1.0 - small_control * standing * upright * move


CHEETAH_RUN_SPEED = 10.0  # dm_control/suite/cheetah.py:36


def cheetah_run_cost(obs):
    """1 - Cheetah.get_reward (dm_control/dm_control/suite/cheetah.py:91-97): linear tolerance of the forward
    speed up to 10 m/s.  physics.speed() is the torso_subtreelinvel sensor (cheetah.py:59-61), which is NOT
    part of the observation (qpos[1:] | qvel, cheetah.py:83-89): the root-x joint velocity obs[8] stands in
    for it -- a documented PROXY, pinned here only through tolerance() itself.
    obs [..., 17] -> cost [...]."""
    obs = np.asarray(obs, dtype=np.float64)
    r = tolerance(obs[..., 8], bounds=(CHEETAH_RUN_SPEED, np.inf), margin=CHEETAH_RUN_SPEED, value_at_margin=0, sigmoid="linear")
    return 1.0 - r


WALKER_STAND_HEIGHT = 1.2  # dm_control/suite/walker.py:37
WALKER_WALK_SPEED = 1.0    # dm_control/suite/walker.py:40


def walker_walk_cost(obs, move_speed=WALKER_WALK_SPEED):
    """1 - PlanarWalker.get_reward (dm_control/dm_control/suite/walker.py:135-158) at move_speed 1 from the
    observation (walker.py:127-133): orientations[0:14] (xx, xz of every body; torso first, and in this
    planar model xx == zz == physics.torso_upright()), height[14] (torso_height), velocity[15:24] (qvel:
    rootz, rootx, rooty, joints).  The horizontal-velocity sensor is not observed: the root-x joint
    velocity obs[16] is the documented PROXY.  obs [..., 24] -> cost [...]."""
    obs = np.asarray(obs, dtype=np.float64)
    standing = tolerance(obs[..., 14], bounds=(WALKER_STAND_HEIGHT, np.inf), margin=WALKER_STAND_HEIGHT / 2)
    upright = (1 + obs[..., 0]) / 2
    stand_reward = (3 * standing + upright) / 4
    move = tolerance(obs[..., 16], bounds=(move_speed, np.inf), margin=move_speed / 2, value_at_margin=0.5, sigmoid="linear")
    return 1.0 - stand_reward * (5 * move + 1) / 6
