"""Plugging the B200 planners into the reference's training loop without editing its files
(SURVEY.md section 8f, rows 1-2).

The reference selects planners through an Enum and a module-level argparse table
(src/mbrl/experiment.py:15-26, :152, :169) and `MPCPolicy` never forwards planner kwargs
(src/mbrl/agents.py:48-55), so two things are needed from the outside:

    import src.mbrl.experiment as experiment
    from mbrl_b200.integration import register_planners, configure
    register_planners(experiment)                       # adds --planner rs-b200 / cem-b200
    configure("cem-b200", num_trajectories=16384, num_iterations=5, elite_frac=0.1)
    experiment.main()                                   # the reference's own entry point

`configure` edits the planner class's `defaults` dict in place -- the same pattern the reference
uses for its own hyper-parameters (src/mbrl/planners.py:141,153-155).  The class objects stay
module-level, so policies holding them still pickle across the reference's worker pool
(src/mbrl/parallel.py:23-38).
"""
from __future__ import annotations

from enum import Enum
from typing import Dict

from .planners import CEMPlanner, GradientDescentPlanner, RandomShootingPlanner

PLANNERS: Dict[str, type] = {"rs-b200": RandomShootingPlanner, "cem-b200": CEMPlanner, "grad-b200": GradientDescentPlanner}
_MEMBER_NAMES = {"rs-b200": "RandomShootingB200", "cem-b200": "CEMB200", "grad-b200": "GradientDescentB200"}


def configure(planner, **hyper):
    """Set default hyper-parameters of a B200 planner (by CLI value or class).  Unknown keys are
    rejected so that a typo does not silently plan with the old value."""
    cls = PLANNERS[planner] if isinstance(planner, str) else planner
    unknown = sorted(set(hyper) - set(cls.defaults))
    if unknown:
        raise KeyError(f"{cls.__name__} has no hyper-parameter(s) {unknown}; known: {sorted(cls.defaults)}")
    cls.defaults.update(hyper)
    return cls


def _add_member(enum_cls, name: str, value):
    """Append a member to an already-built Enum (what `aenum.extend_enum` does)."""
    if value in enum_cls._value2member_map_:
        return enum_cls._value2member_map_[value]
    if name in enum_cls._member_map_:
        raise ValueError(f"{enum_cls.__name__}.{name} already exists with another value")
    member = object.__new__(enum_cls)
    member._name_ = name
    member._value_ = value
    if hasattr(member, "_sort_order_"):
        member._sort_order_ = len(enum_cls._member_names_)
    enum_cls._member_names_.append(name)
    enum_cls._member_map_[name] = member
    enum_cls._value2member_map_[value] = member
    type.__setattr__(enum_cls, name, member)
    return member


def register_planners(experiment_module, enum_name: str = "Planner"):
    """Adds `rs-b200` / `cem-b200` to the reference's planner Enum, teaches `construct()` to return
    the B200 classes and refreshes every argparse table entry whose `type` is that Enum
    (experiment.py:152 captured `list(Planner)` at import time).  Idempotent."""
    enum_cls = getattr(experiment_module, enum_name)
    if not (isinstance(enum_cls, type) and issubclass(enum_cls, Enum)):
        raise TypeError(f"{experiment_module.__name__}.{enum_name} is not an Enum")
    members = {value: _add_member(enum_cls, _MEMBER_NAMES[value], value) for value in PLANNERS}
    # MPCPolicy consumes only plan[1][0], the first action (src/mbrl/agents.py:56): planners selected
    # through the reference's CLI skip the fp32 replay that produces the predicted states (zeros are
    # returned in their place; the next call's warm start ignores the states anyway) and CEM keeps its
    # sampling mean resident on the device between MPC steps.  configure(..., return_states=True) undoes it.
    RandomShootingPlanner.defaults["return_states"] = False
    CEMPlanner.defaults["return_states"] = False
    CEMPlanner.defaults["warm_start"] = "shift_mean"
    if not getattr(enum_cls.construct, "_b200_patched", False):
        original = enum_cls.construct

        def construct(self, *args, **kwargs):
            if self.value in PLANNERS:
                return PLANNERS[self.value]
            return original(self, *args, **kwargs)

        construct._b200_patched = True
        type.__setattr__(enum_cls, "construct", construct)
    for obj in vars(experiment_module).values():
        entries = obj if isinstance(obj, (list, tuple)) else ()
        for entry in entries:
            if isinstance(entry, dict) and entry.get("type") is enum_cls and "choices" in entry:
                entry["choices"] = list(enum_cls)
    return members
