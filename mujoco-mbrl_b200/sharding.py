"""Multi-GPU planning: one process per GPU, torch.distributed for the plumbing.

* Population sharding (BASELINE config 4): rank r owns global candidates
  [r*N_local, (r+1)*N_local).  Philox counters carry the GLOBAL candidate index, so the
  population is identical for any number of GPUs.  One exchange step per CEM iteration: an
  all-gather of each rank's k_local cheapest (cost, global index) pairs (8*k_local bytes per
  rank; NCCL over NVLink).  Every rank then selects the same global top-k and refits
  redundantly by regenerating the elite actions from their global indices -- the refit needs
  no second collective and is bit-identical on all ranks (the north_star's "broadcast the
  refit mean/std" is available as `broadcast_refit=True`, e.g. to assert that equality).
  This module is the readable, gloo-tested statement of that host logic.  The production loop
  runs inside the library (native.NativePlanner.p2p_init / comm_init + plan): same elites, but
  over peer memory the refit is distributed -- per-rank partial sums added in rank order -- so
  its mean/std equal this module's up to fp32 rounding (and exactly at world size 1).
* Environment sharding (config 5): independent environments, no collective at all --
  `env_shard()` just computes each rank's slice; the caller runs a plain NativePlanner on it.

The device work goes through an `ops` object (the native C-ABI by default).  CPU gloo tests
bind a stand-in built on the oracle to exercise exactly this host logic without a GPU.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def env_shard(num_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """(first env, count) of rank's contiguous slice; remainders go to the low ranks."""
    base, rem = divmod(num_envs, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


class NativeOps:
    """Device ops backed by libmbrl_b200.so (the product path)."""

    def __init__(self, planner):
        from . import native
        self.native = native
        self.h = planner

    def rollout(self, d_s0, seed, it, d_mu, d_sd, cand_offset):
        return self.h.rollout(d_s0, self.native.SAMPLE_GAUSSIAN, seed, it, d_mu=d_mu, d_sd=d_sd,
                              cand_offset=cand_offset)[0]

    def topk(self, d_costs, k):
        idx, cost, best = self.native.topk(d_costs, k, 1)
        return idx[0], cost[0]

    def refit(self, d_elite_global, k, seed, it, d_mu, d_sd):
        return self.h.refit(d_elite_global.view(1, -1), k, self.native.SAMPLE_GAUSSIAN, seed, it, d_mu=d_mu, d_sd=d_sd,
                            cand_offset=0)

    def emit(self, d_s0, d_best, d_mu_hist, d_sd_hist, iterations, seed):
        return self.h.emit(d_s0, d_best, d_mu_hist, d_sd_hist, iterations, self.native.SAMPLE_GAUSSIAN, seed,
                           cand_offset=0)


class PopulationShardedCEM:
    """CEM over a population split across the ranks of a torch.distributed group."""

    def __init__(self, ops, n_local: int, horizon: int, act_dim: int, rank: int, world: int, group=None,
                 lo: float = -1.0, hi: float = 1.0, broadcast_refit: bool = False):
        self.ops, self.n_local, self.H, self.A = ops, n_local, horizon, act_dim
        self.rank, self.world, self.group = rank, world, group
        self.lo, self.hi = lo, hi
        self.broadcast_refit = broadcast_refit
        self.n_total = n_local * world

    def merge_elites(self, local_cost, local_gidx, k: int):
        """All-gather the per-shard elites and select the global k cheapest.  Gathered in rank
        order with ascending local index inside a rank == ascending GLOBAL index, so the
        selection's tie-break (lower position) is the global lower-index rule."""
        import torch
        import torch.distributed as dist
        packed = torch.stack([local_cost.float(), local_gidx.int().view(torch.float32)])  # [2, k_local]
        if self.world > 1:
            flat = torch.empty(self.world * 2, packed.shape[1], dtype=packed.dtype, device=packed.device)
            dist.all_gather_into_tensor(flat, packed.contiguous(), group=self.group)
            out = flat.view(self.world, 2, -1)
        else:
            out = packed[None]
        costs = out[:, 0].reshape(-1).contiguous()
        gidx = out[:, 1].contiguous().view(torch.int32).reshape(-1)
        pos, pcost = self.ops.topk(costs, k)
        return gidx[pos.long()].contiguous(), pcost

    def plan(self, d_s0, iterations: int, k: int, seed: int = 0):
        """Returns dict(states [1,H,O], actions [1,H,A], best (cost, iteration, global index), mu, sd)."""
        import torch
        import torch.distributed as dist
        dev = d_s0.device
        mu_hist = torch.empty(iterations + 1, 1, self.H, self.A, dtype=torch.float32, device=dev)
        sd_hist = torch.empty_like(mu_hist)
        mu_hist[0].fill_(0.5 * (self.lo + self.hi))
        sd_hist[0].fill_(0.5 * (self.hi - self.lo))
        k_local = min(k, self.n_local)
        best_cost = torch.full((), float("inf"), device=dev)
        best_i = torch.zeros(1, 4, dtype=torch.int32, device=dev)  # MbrlPlanInfo: cost bits, iteration, index, 0
        offset = self.rank * self.n_local
        for it in range(iterations):
            costs = self.ops.rollout(d_s0, seed, it, mu_hist[it], sd_hist[it], offset)
            lidx, lcost = self.ops.topk(costs, k_local)
            elite, ecost = self.merge_elites(lcost, lidx + offset, k)
            # best-ever: strict improvement only (earliest iteration wins ties); among equal
            # costs the lowest global index (elite is in ascending index order)
            cmin = ecost.min()
            first = elite[(ecost == cmin).nonzero()[0, 0]]
            better = cmin < best_cost
            cand = torch.stack([cmin.view(1).view(torch.int32)[0], torch.tensor(it, dtype=torch.int32, device=dev),
                                first.int(), torch.tensor(0, dtype=torch.int32, device=dev)]).view(1, 4)
            best_i = torch.where(better, cand, best_i)
            best_cost = torch.where(better, cmin, best_cost)
            mu, sd = self.ops.refit(elite, k, seed, it, mu_hist[it], sd_hist[it])
            if self.broadcast_refit and self.world > 1:
                dist.broadcast(mu, src=0, group=self.group)
                dist.broadcast(sd, src=0, group=self.group)
            mu_hist[it + 1].copy_(mu.view(1, self.H, self.A))
            sd_hist[it + 1].copy_(sd.view(1, self.H, self.A))
        states, actions = self.ops.emit(d_s0, best_i, mu_hist, sd_hist, iterations, seed)
        return dict(states=states, actions=actions, best=best_i, mu=mu_hist[iterations], sd=sd_hist[iterations])
