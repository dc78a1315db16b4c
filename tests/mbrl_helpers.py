"""Shared helpers for the test-suite (oracle side; imported as a top-level module because
pytest puts this directory on sys.path)."""
import os

import numpy as np
import torch

from oracle.planner_oracle import PlannerParams

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def params_from_golden(g) -> PlannerParams:
    t = lambda k: torch.from_numpy(np.ascontiguousarray(g[k])).float()
    if "W2" not in g:  # LinearModel fixture: a single Linear(D, O)
        return PlannerParams(
            t("W1"), t("b1"), None, None, None, None,
            t("mu_s"), t("sd_s"), t("mu_a"), t("sd_a"), t("cost_w"), t("goal"),
            alpha=float(g["alpha"]), beta=float(g["beta"]), act_lo=float(g["lo"]), act_hi=float(g["hi"]),
        )
    if "W4" in g:  # ModelWithReward fixture: no SmoothAbs/Cosh parameters, a reward head instead
        obs = g["W3"].shape[0]
        return PlannerParams(
            t("W1"), t("b1"), t("W2"), t("b2"), t("W3"), t("b3"),
            t("mu_s"), t("sd_s"), t("mu_a"), t("sd_a"), torch.ones(obs), torch.zeros(obs),
            act_lo=float(g["lo"]), act_hi=float(g["hi"]),
            W4=t("W4"), b4=t("b4"), mu_r=float(g["mu_r"][0]), sd_r=float(g["sd_r"][0]),
        )
    return PlannerParams(
        t("W1"), t("b1"), t("W2"), t("b2"), t("W3"), t("b3"),
        t("mu_s"), t("sd_s"), t("mu_a"), t("sd_a"), t("cost_w"), t("goal"),
        alpha=float(g["alpha"]), beta=float(g["beta"]),
        act_lo=float(g["lo"]), act_hi=float(g["hi"]),
    )
