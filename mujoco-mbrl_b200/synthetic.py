"""Synthetic planning problems of SURVEY.md section 8(d) (benchmark / smoke inputs):
default nn.Linear init (the initialiser Model.__init__ gets, src/mbrl/models.py:99-101),
mu_s~N(0,1), sd_s~U(0.5,1.5), mu_a=0, sd_a=1/sqrt(3), SmoothAbs w=1 g=0 alpha=0.4,
Cosh beta=0.25, action bounds +-1; s0 ~ N(mu_s, sd_s) with seed 1000+call."""
import math

import torch

from .adaptor import PlanningProblem


def synthetic_problem(obs_dim: int, act_dim: int, hidden: int, seed: int = 0) -> PlanningProblem:
    g = torch.Generator().manual_seed(seed)

    def linear(out_f, in_f):
        bound = 1.0 / math.sqrt(in_f)
        W = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
        b = (torch.rand(out_f, generator=g) * 2 - 1) * bound
        return W, b

    W1, b1 = linear(hidden, obs_dim + act_dim)
    W2, b2 = linear(hidden, hidden)
    W3, b3 = linear(obs_dim, hidden)
    mu_s = torch.randn(obs_dim, generator=g)
    sd_s = torch.rand(obs_dim, generator=g) + 0.5
    return PlanningProblem(W1, b1, W2, b2, W3, b3, mu_s, sd_s, torch.zeros(act_dim),
                           torch.full((act_dim,), 1.0 / math.sqrt(3.0)), torch.ones(obs_dim), torch.zeros(obs_dim))


def synthetic_state(prob: PlanningProblem, call: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(1000 + call)
    return prob.mu_s + prob.sd_s * torch.randn(prob.obs_dim, generator=g)
