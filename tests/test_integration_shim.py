"""CPU tests of the registration shim: an Enum + argparse table shaped like the
reference's src/mbrl/experiment.py:15-26,152 gets the B200 planners added without source edits;
when /root/reference is present (this container only) the same is done to the real module."""
import argparse
import os
import pickle
import sys
import types
from enum import Enum

import pytest


@pytest.fixture(autouse=True)
def _restore_planner_defaults():
    """register_planners / configure edit the planner classes' `defaults` in place (the reference's own
    hyper-parameter pattern); keep the edits from leaking into other tests of the session."""
    from mbrl_b200 import CEMPlanner, RandomShootingPlanner
    saved = dict(CEMPlanner.defaults), dict(RandomShootingPlanner.defaults)
    yield
    CEMPlanner.defaults.clear(); CEMPlanner.defaults.update(saved[0])
    RandomShootingPlanner.defaults.clear(); RandomShootingPlanner.defaults.update(saved[1])


def _replica_module():
    mod = types.ModuleType("replica_experiment")

    class Planner(Enum):
        RandomShooting = "rs"
        GradientDescent = "grad"

        def __str__(self):
            return self.value

        def construct(self):
            return {"rs": "ref-rs", "grad": "ref-grad"}[self.value]

    mod.Planner = Planner
    mod.CONFIG = [{"name": "planner", "type": Planner, "choices": list(Planner)},
                  {"name": "horizon", "type": int, "default": 20}]
    return mod


def _parse(mod, argv):
    parser = argparse.ArgumentParser()
    for entry in mod.CONFIG:
        e = dict(entry)
        parser.add_argument("--" + e.pop("name"), **e)
    return parser.parse_args(argv)


def test_register_on_replica_enum():
    from mbrl_b200 import CEMPlanner, RandomShootingPlanner
    from mbrl_b200.integration import register_planners
    mod = _replica_module()
    register_planners(mod)
    register_planners(mod)  # idempotent
    assert [str(m) for m in mod.Planner] == ["rs", "grad", "rs-b200", "cem-b200", "grad-b200"]
    assert mod.Planner("cem-b200").construct() is CEMPlanner
    assert mod.Planner("rs-b200").construct() is RandomShootingPlanner
    assert mod.Planner("rs").construct() == "ref-rs"          # reference members untouched
    # MPC use (MPCPolicy reads only the first action): no state replay, device-resident CEM warm start
    assert CEMPlanner.defaults["return_states"] is False and RandomShootingPlanner.defaults["return_states"] is False
    assert CEMPlanner.defaults["warm_start"] == "shift_mean"
    assert _parse(mod, ["--planner", "cem-b200"]).planner is mod.Planner.CEMB200
    with pytest.raises(SystemExit):
        _parse(mod, ["--planner", "nope"])


def test_configure_sets_defaults_and_rejects_typos():
    from mbrl_b200 import CEMPlanner
    from mbrl_b200.integration import configure
    saved = dict(CEMPlanner.defaults)
    try:
        assert configure("cem-b200", num_trajectories=4096, num_iterations=3) is CEMPlanner
        assert CEMPlanner.defaults["num_trajectories"] == 4096 and CEMPlanner.defaults["num_iterations"] == 3
        with pytest.raises(KeyError):
            configure(CEMPlanner, num_trajectory=1)
        assert pickle.loads(pickle.dumps(CEMPlanner)) is CEMPlanner  # class object still pickles by name
    finally:
        CEMPlanner.defaults.clear()
        CEMPlanner.defaults.update(saved)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/mbrl"), reason="reference tree not present on this box")
def test_register_on_reference_experiment_module():
    from mbrl_b200 import CEMPlanner
    from mbrl_b200.integration import register_planners
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden  # noqa: F401  (installs the import stubs for tensorboardX / dm_control / colorlog)
    make_golden._install_stubs()
    sys.path.insert(0, "/root/reference")
    import src.mbrl.experiment as experiment
    register_planners(experiment)
    assert experiment.Planner("cem-b200").construct() is CEMPlanner
    assert experiment.Planner("rs").construct().__name__ == "RandomShootingPlanner"
    entry = [e for e in experiment.CONFIG_DEF if e.get("name") == "planner"]  # experiment.py:152
    assert entry and experiment.Planner.CEMB200 in entry[0]["choices"]
