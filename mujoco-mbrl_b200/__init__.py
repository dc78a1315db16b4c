"""mbrl_b200 -- B200-native MPC planning hot path for Khodeir/mujoco-mbrl.

Drop-in for ``src/mbrl/planners.py``: ``RandomShootingPlanner`` / ``CEMPlanner`` keep the
reference's static ``plan(initial_state, model, cost, sample_action, horizon,
initial_trajectory=None, **kwargs) -> (states, actions)`` API and run the candidate
rollout, cost, elite selection and refit in hand-written sm_100a CUDA behind a C ABI
(``include/mbrl_b200.h``).  No CPU fallback: importing the planners works anywhere,
planning requires the built library and a CUDA device.
"""
from .adaptor import PlanningProblem, problem_from_callables  # noqa: F401
from .native import NativePlanner, load_library  # noqa: F401
from .planners import CEMPlanner, GradientDescentPlanner, ModelPlanner, RandomShootingPlanner  # noqa: F401
from .integration import configure, register_planners  # noqa: F401

__all__ = [
    "CEMPlanner", "GradientDescentPlanner", "ModelPlanner", "RandomShootingPlanner", "NativePlanner", "PlanningProblem",
    "problem_from_callables", "load_library", "configure", "register_planners",
]
