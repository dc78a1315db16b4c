"""GPU tests of the sharding invariants (single GPU: the ranks' shards are run back to back on
one device, never as concurrently waiting kernels):
  * population sharding: Philox counters carry the GLOBAL candidate index, so the costs of two
    half-population shards are bit-identical to the unsharded rollout, and the redundant refit
    from global elite indices equals the unsharded refit;
  * the host loop PopulationShardedCEM (world 1) reproduces NativePlanner.plan;
  * environment sharding: env_offset makes a shard of environments draw the same streams."""
import os
import numpy as np
import pytest
import torch

from oracle import planner_oracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def native():
    from mbrl_b200 import native as n
    assert torch.cuda.is_available()
    return n


def _planner(native, p, H, n, envs=1, iters=1, engine="fp32"):
    h = native.NativePlanner(p.obs_dim, p.act_dim, p.hidden, H, n, envs, iters, None, engine)
    h.set_weights(p.W1, p.b1, p.W2, p.b2, p.W3, p.b3)
    h.set_norm(p.mu_s, p.sd_s, p.mu_a, p.sd_a)
    h.set_cost(p.cost_w, p.goal, p.alpha, p.beta)
    h.set_action_bounds(p.act_lo, p.act_hi)
    return h


@pytest.mark.parametrize("engine", ["fp32", "fp16"])
def test_population_shards_are_bit_identical_to_unsharded(native, engine):
    p = po.synthetic_params(24, 6, 200)  # walker-walk shape (BASELINE config 4)
    H, N, k = 30, 2048, 204
    s0 = po.synthetic_state(p, 0)[None].cuda()
    mu = (torch.rand(1, H, 6) * 0.2 - 0.1).cuda()
    sd = (torch.rand(1, H, 6) * 0.5 + 0.5).cuda()
    full = _planner(native, p, H, N, engine=engine)
    half = _planner(native, p, H, N // 2, engine=engine)
    c_full, _, _ = full.rollout(s0, native.SAMPLE_GAUSSIAN, 77, 3, d_mu=mu, d_sd=sd)
    c0, _, _ = half.rollout(s0, native.SAMPLE_GAUSSIAN, 77, 3, d_mu=mu, d_sd=sd, cand_offset=0)
    c1, _, _ = half.rollout(s0, native.SAMPLE_GAUSSIAN, 77, 3, d_mu=mu, d_sd=sd, cand_offset=N // 2)
    assert torch.equal(c_full, torch.cat([c0, c1]))
    # per-shard elites -> merged global elites == unsharded elites
    idx_full, _, _ = native.topk(c_full, k, 1)
    i0, e0, _ = native.topk(c0, k, 1)
    i1, e1, _ = native.topk(c1, k, 1)
    gathered_cost = torch.cat([e0[0], e1[0]])
    gathered_idx = torch.cat([i0[0], i1[0] + N // 2])
    pos, _, _ = native.topk(gathered_cost.contiguous(), k, 1)
    assert torch.equal(gathered_idx[pos[0].long()], idx_full[0])
    # redundant refit from global indices on a shard-sized handle == unsharded refit
    m_full, s_full = full.refit(idx_full, k, native.SAMPLE_GAUSSIAN, 77, 3, d_mu=mu, d_sd=sd)
    m_half, s_half = half.refit(idx_full, k, native.SAMPLE_GAUSSIAN, 77, 3, d_mu=mu, d_sd=sd, cand_offset=0)
    assert torch.equal(m_full, m_half) and torch.equal(s_full, s_half)


def test_sharded_host_loop_world1_matches_plan(native):
    from mbrl_b200.sharding import NativeOps, PopulationShardedCEM
    p = po.synthetic_params(17, 6, 200)
    H, N, I, k = 30, 4096, 4, 409
    s0 = po.synthetic_state(p, 2)
    h = _planner(native, p, H, N, 1, I, engine="fp16")
    out = h.plan(s0.numpy(), I, k, native.SAMPLE_GAUSSIAN, seed=9, want_dist=True)
    cem = PopulationShardedCEM(NativeOps(h), N, H, 6, 0, 1)
    res = cem.plan(s0[None].cuda(), I, k, seed=9)
    best = res["best"].cpu().numpy()[0]
    assert (int(best[1]), int(best[2])) == (int(out["info"]["best_iteration"][0]), int(out["info"]["best_index"][0]))
    assert best[0:1].view(np.float32)[0] == out["info"]["best_cost"][0]
    np.testing.assert_array_equal(res["actions"].cpu().numpy(), out["actions"])
    np.testing.assert_array_equal(res["states"].cpu().numpy(), out["states"])
    np.testing.assert_array_equal(res["mu"].cpu().numpy().reshape(out["mu"].shape), out["mu"])


def test_environment_shards_draw_the_same_streams(native):
    p = po.synthetic_params(9, 3, 40)
    H, N, E = 6, 300, 4
    s0 = torch.stack([po.synthetic_state(p, e) for e in range(E)]).cuda()
    mu = torch.zeros(E, H, 3, device="cuda")
    sd = torch.ones(E, H, 3, device="cuda")
    allenv = _planner(native, p, H, N, E)
    two = _planner(native, p, H, N, 2)
    c_all, _, _ = allenv.rollout(s0, native.SAMPLE_GAUSSIAN, 5, 1, d_mu=mu, d_sd=sd)
    c_hi, _, _ = two.rollout(s0[2:].contiguous(), native.SAMPLE_GAUSSIAN, 5, 1, d_mu=mu[2:].contiguous(),
                             d_sd=sd[2:].contiguous(), env_offset=2)
    assert torch.equal(c_all[2 * N:], c_hi)


def test_in_library_nccl_world1_equals_plain_plan(native):
    """mbrl_comm_init with a 1-rank communicator: the sharded code path (pack, ncclAllGather,
    merge, remap, global-index refit) must reproduce the plain plan bit for bit."""
    p = po.synthetic_params(17, 6, 200)
    H, N, I, k = 30, 4096, 4, 409
    s0 = po.synthetic_state(p, 7).numpy()
    plain = _planner(native, p, H, N, 1, I, engine="fp16")
    want = plain.plan(s0, I, k, native.SAMPLE_GAUSSIAN, seed=3, want_dist=True)
    sh = _planner(native, p, H, N, 1, I, engine="fp16")
    sh.comm_init(0, 1)
    got = sh.plan(s0, I, k, native.SAMPLE_GAUSSIAN, seed=3, want_dist=True)
    for key in ("actions", "states", "mu", "sd"):
        np.testing.assert_array_equal(got[key], want[key])
    for key in ("best_cost", "best_index", "best_iteration"):
        np.testing.assert_array_equal(got["info"][key], want["info"][key])


def test_refit_segments_change_only_the_rounding(native):
    """mbrl_set_refit_segments(W): the refit adds W per-segment partial sums in segment order -- the
    arithmetic of a W-way sharded plan (tests/multi_gpu_check.py proves that identity on >= 2 GPUs).
    Against the plain order the mean / std move by fp32 rounding only (tolerance 2e-6 absolute on
    values in [-1, 1]), one segment IS the plain order, and the setter validates its argument."""
    p = po.synthetic_params(17, 6, 200)
    H, N, I, k = 30, 8192, 3, 819
    s0 = po.synthetic_state(p, 5).numpy()
    plain = _planner(native, p, H, N, 1, I, engine="fp32")
    want = plain.plan(s0, I, k, native.SAMPLE_GAUSSIAN, seed=9, want_dist=True)
    seg = _planner(native, p, H, N, 1, I, engine="fp32")
    seg.set_refit_segments(1)
    one = seg.plan(s0, I, k, native.SAMPLE_GAUSSIAN, seed=9, want_dist=True)
    for key in ("actions", "mu", "sd"):
        np.testing.assert_array_equal(one[key], want[key])
    for w in (2, 8, 64):
        seg.set_refit_segments(w)
        got = seg.plan(s0, 1, k, native.SAMPLE_GAUSSIAN, seed=9, want_dist=True)  # one iteration: same elites for sure
        ref = plain.plan(s0, 1, k, native.SAMPLE_GAUSSIAN, seed=9, want_dist=True)
        np.testing.assert_allclose(got["mu"], ref["mu"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(got["sd"], ref["sd"], rtol=0, atol=2e-6)
        assert got["info"]["best_index"][0] == ref["info"]["best_index"][0]
    with pytest.raises(RuntimeError):
        seg.set_refit_segments(3)      # does not divide the population
    with pytest.raises(RuntimeError):
        seg.set_refit_segments(128)    # more than 64


def test_multi_gpu_population_sharding_is_bit_identical():
    """Runs tests/multi_gpu_check.py under torchrun when this box has >= 2 GPUs (skipped on the
    single-GPU test box): every rank's plan == the unsharded plan, both transports, incl. the full
    per-GPU size of BASELINE config 4 with a chunked refit."""
    import subprocess
    import sys
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multi_gpu_check.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", script]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("multi_gpu_check[") == 11, res.stdout[-2000:]
