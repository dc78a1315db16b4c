"""Multi-GPU check of the in-library NCCL population sharding (run under torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py

Every rank plans on its shard; all ranks must return the same plan, and it must equal the
UNSHARDED plan over the whole population (bit for bit: Philox counters carry global indices,
the merge keeps the lower-global-index tie rule, the refit regenerates elites from global indices and
adds per-rank partial sums in rank order -- the unsharded plan is asked for the same order with
set_refit_segments(world))."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mbrl_b200 import native  # noqa: E402
from mbrl_b200.synthetic import synthetic_problem, synthetic_state  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ["MBRL_P2P_TIMEOUT_S"] = "2"   # for the last case; microseconds are what every other case needs
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    O, A, U, H, I = 24, 6, 200, 30, 4  # walker-walk shape (BASELINE config 4)
    prob = synthetic_problem(O, A, U)
    s0 = synthetic_state(prob, 4).numpy()
    # the last case is the full per-GPU size of BASELINE config 4 (16384 candidates per rank): the global
    # elite set exceeds 2048, so the refit runs as chunk CTAs whose partial sums must add up in the same
    # order on every rank and in the unsharded plan
    for engine, force_kl, transport, n_local in (("fp32", None, "nccl", 2048), ("fp16", None, "nccl", 2048),
                                                 ("fp16", "min", "nccl", 2048), ("fp16", None, "p2p", 2048),
                                                 ("fp16", "min", "p2p", 2048), ("fp16", None, "p2p", 16384)):
        n_total, k = n_local * world, int(0.1 * n_local * world)
        if force_kl is None:
            os.environ.pop("MBRL_SHARD_KL", None)
        else:  # a gather that is far too small: the on-device check must flag it and the plan is redone in full
            os.environ["MBRL_SHARD_KL"] = force_kl
        h = native.NativePlanner(O, A, U, H, n_local, 1, I, k, engine, local)
        h.load_problem(prob)
        if transport == "p2p":
            assert h.p2p_init(rank, world)   # NVLink peer stores + sequence flags instead of ncclAllGather
        else:
            h.comm_init(rank, world)
        for rep in range(3):          # repeated plans: sequence numbers / double buffering keep working
            out = h.plan(s0, I, k, native.SAMPLE_GAUSSIAN, seed=20 + rep, want_dist=True)
        out = h.plan(s0, I, k, native.SAMPLE_GAUSSIAN, seed=21, want_dist=True)
        mine = torch.from_numpy(np.concatenate([out["actions"].ravel(), out["states"].ravel(), out["mu"].ravel(),
                                                out["sd"].ravel(), out["info"]["best_cost"],
                                                out["info"]["best_index"].astype(np.float32),
                                                out["info"]["best_iteration"].astype(np.float32)])).cuda()
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(mine, ref), f"rank {rank}: plan differs from rank 0 ({engine})"
        if rank == 0:
            full = native.NativePlanner(O, A, U, H, n_total, 1, I, k, engine, local)
            full.load_problem(prob)
            full.set_refit_segments(world)  # the sharded refit adds per-rank partial sums in rank order
            want = full.plan(s0, I, k, native.SAMPLE_GAUSSIAN, seed=21, want_dist=True)
            for key in ("actions", "states", "mu", "sd"):
                np.testing.assert_array_equal(out[key], want[key], err_msg=f"{engine} {key}")
            for key in ("best_cost", "best_index", "best_iteration"):
                np.testing.assert_array_equal(out["info"][key], want["info"][key], err_msg=f"{engine} {key}")
            print(f"multi_gpu_check[{engine}, {transport}{', forced tiny gather' if force_kl else ''}]: {world} ranks == unsharded N={n_total}: best cost "
                  f"{out['info']['best_cost'][0]:.4f} idx {out['info']['best_index'][0]} it {out['info']['best_iteration'][0]}")
        dist.barrier()
        h.close()
    # random shooting (I = 1, k = 1: the reference's planner, src/mbrl/planners.py:166-187) over a sharded
    # population: one (cost, index) pair per rank is exchanged and the global argmin (ties -> lower global
    # index, as np.argmin) must be the unsharded one; uniform sampler as EnvWrapper._sample_action
    for engine, transport in (("fp32", "nccl"), ("fp16", "p2p")):
        n_local = 4096
        h = native.NativePlanner(O, A, U, H, n_local, 1, 1, 1, engine, local)
        h.load_problem(prob)
        if transport == "p2p":
            assert h.p2p_init(rank, world)
        else:
            h.comm_init(rank, world)
        for rep in range(2):
            out = h.plan(s0, 1, 1, native.SAMPLE_UNIFORM, seed=30 + rep)
        mine = torch.from_numpy(np.concatenate([out["actions"].ravel(), out["states"].ravel(), out["info"]["best_cost"],
                                                out["info"]["best_index"].astype(np.float32)])).cuda()
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(mine, ref), f"rank {rank}: RS plan differs from rank 0 ({engine})"
        if rank == 0:
            full = native.NativePlanner(O, A, U, H, n_local * world, 1, 1, 1, engine, local)
            full.load_problem(prob)
            want = full.plan(s0, 1, 1, native.SAMPLE_UNIFORM, seed=31)
            for key in ("actions", "states"):
                np.testing.assert_array_equal(out[key], want[key], err_msg=f"RS {engine} {key}")
            for key in ("best_cost", "best_index"):
                np.testing.assert_array_equal(out["info"][key], want["info"][key], err_msg=f"RS {engine} {key}")
            print(f"multi_gpu_check[RS, {engine}, {transport}]: {world} ranks == unsharded N={n_local * world}: "
                  f"argmin {out['info']['best_index'][0]} cost {out['info']['best_cost'][0]:.4f}")
        dist.barrier()
        h.close()
    # Every candidate costs the same (zero state weights, a flat action term): the elite set is decided by
    # the tie rule alone -- the k lowest GLOBAL indices, all of them on rank 0.  Exercises ties straddling
    # the cut inside the merge's own-slice compaction, ranks without a single elite (an empty partial sum),
    # and the exactness flag + full-size redo (a rank's threshold equals the global one).
    import dataclasses
    flat = dataclasses.replace(prob, cost_w=torch.zeros(O), beta=1e30)
    for transport in ("nccl", "p2p"):
        n_local, k = 2048, int(0.1 * 2048 * world)
        h = native.NativePlanner(O, A, U, H, n_local, 1, I, k, "fp16", local)
        h.load_problem(flat)
        if transport == "p2p":
            assert h.p2p_init(rank, world)
        else:
            h.comm_init(rank, world)
        out = h.plan(s0, I, k, native.SAMPLE_GAUSSIAN, seed=50, want_dist=True)
        mine = torch.from_numpy(np.concatenate([out["actions"].ravel(), out["mu"].ravel(), out["sd"].ravel(),
                                                out["info"]["best_index"].astype(np.float32)])).cuda()
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(mine, ref), f"rank {rank}: all-ties plan differs from rank 0 ({transport})"
        if rank == 0:
            full = native.NativePlanner(O, A, U, H, n_local * world, 1, I, k, "fp16", local)
            full.load_problem(flat)
            full.set_refit_segments(world)
            want = full.plan(s0, I, k, native.SAMPLE_GAUSSIAN, seed=50, want_dist=True)
            for key in ("actions", "states", "mu", "sd"):
                np.testing.assert_array_equal(out[key], want[key], err_msg=f"ties {transport} {key}")
            assert out["info"]["best_index"][0] == want["info"]["best_index"][0] == 0
            print(f"multi_gpu_check[all costs tie, {transport}]: {world} ranks == unsharded, elites = the {k} lowest global indices")
        dist.barrier()
        h.close()
    # A rank that never shows up: the survivor must get an error (deterministic sentinels inside, never a
    # plausible plan built from the previous iteration's packets).  MBRL_P2P_TIMEOUT_S is read once per
    # process, before the first sharded plan -- main() set it to 2 s.
    h = native.NativePlanner(O, A, U, H, 2048, 1, 2, 204, "fp16", local)
    h.load_problem(prob)
    assert h.p2p_init(rank, world)
    h.plan(s0, 2, 204, native.SAMPLE_GAUSSIAN, seed=40)   # everybody: fine
    dist.barrier()
    if rank == 0:
        try:
            h.plan(s0, 1, 204, native.SAMPLE_GAUSSIAN, seed=41, want_dist=True)  # alone
            raise AssertionError("a plan whose peers never sent anything returned normally")
        except native.MbrlError as e:
            assert "timed out" in str(e), str(e)
            print(f"multi_gpu_check[timeout]: rank 0 alone -> {str(e)[:60]}...")
    dist.barrier()
    h.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
