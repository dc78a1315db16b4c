"""world_size-2 gloo test (CPU) of the population-sharding host logic: the elite all-gather,
the global re-selection with the lower-global-index tie rule, the redundant refit and the
best-ever tracking must give the same plan for 1 rank and for 2 ranks.  Device ops are bound
to an oracle-backed stand-in (TEST INFRASTRUCTURE) -- the product binds the C-ABI library."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import philox as ophilox
from oracle import planner_oracle as po


class OracleOps:
    def __init__(self, p, horizon, n_local):
        self.p, self.H, self.n = p, horizon, n_local

    def _actions(self, seed, it, mu, sd, n, offset):
        z = torch.from_numpy(ophilox.standard_normal(seed, it, self.H, n, self.p.act_dim, cand_offset=offset))
        return po.gaussian_actions(mu.view(self.H, -1), sd.view(self.H, -1), z, n, self.p.act_lo, self.p.act_hi)

    def rollout(self, d_s0, seed, it, d_mu, d_sd, cand_offset):
        acts = self._actions(seed, it, d_mu, d_sd, self.n, cand_offset)
        _, c = po.rollout_costs(self.p, d_s0[0], acts, self.H, self.n)
        # quantise so that exact ties across shards occur and the tie rule is exercised
        return torch.from_numpy(np.round(c * 4) / 4).float()

    def topk(self, d_costs, k):
        idx = np.sort(po.topk_stable(d_costs.numpy(), k))
        return torch.from_numpy(idx).int(), d_costs[torch.from_numpy(idx).long()]

    def refit(self, d_elite_global, k, seed, it, d_mu, d_sd):
        cols = []
        for gi in d_elite_global.tolist():  # regenerate each elite from its GLOBAL index
            cols.append(self._actions(seed, it, d_mu, d_sd, 1, gi).view(self.H, 1, -1))
        a = torch.cat(cols, dim=1)
        return a.mean(1)[None], a.std(1, unbiased=False)[None]

    def emit(self, d_s0, d_best, d_mu_hist, d_sd_hist, iterations, seed):
        _, it, gi, _ = d_best[0].tolist()
        acts = self._actions(seed, it, d_mu_hist[it], d_sd_hist[it], 1, gi)
        st, _ = po.rollout_costs(self.p, d_s0[0], acts, self.H, 1)
        return st.view(1, self.H, -1), acts.view(1, self.H, -1)


def _run(rank, world, port, n_total, out):
    from mbrl_b200.sharding import PopulationShardedCEM
    torch.set_num_threads(1)
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
    p = po.synthetic_params(5, 2, 16, seed=2)
    H, k, iters = 6, 24, 3
    n_local = n_total // world
    cem = PopulationShardedCEM(OracleOps(p, H, n_local), n_local, H, p.act_dim, rank, world, broadcast_refit=(world > 1))
    res = cem.plan(po.synthetic_state(p, 1)[None], iters, k, seed=5)
    out[(world, rank)] = {k_: v.numpy().copy() for k_, v in res.items()}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_population_sharding_world2_matches_world1():
    mgr = mp.Manager()
    out = mgr.dict()
    _run(0, 1, 0, 256, out)
    port = _free_port()
    mp.spawn(_run, args=(2, port, 256, out), nprocs=2, join=True)
    one, r0, r1 = out[(1, 0)], out[(2, 0)], out[(2, 1)]
    for key in ("states", "actions", "best", "mu", "sd"):
        np.testing.assert_array_equal(r0[key], r1[key])       # ranks agree bit-for-bit
        np.testing.assert_array_equal(one[key], r0[key])      # and with the unsharded run
    assert one["best"][0, 1] >= 0 and 0 <= one["best"][0, 2] < 256


def test_env_shard_partitions_everything():
    from mbrl_b200.sharding import env_shard
    for n, w in [(1024, 8), (10, 4), (3, 8), (128, 1)]:
        spans = [env_shard(n, r, w) for r in range(w)]
        assert sum(c for _, c in spans) == n
        nxt = 0
        for first, c in spans:
            assert first == nxt
            nxt += c
