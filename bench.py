#!/usr/bin/env python
"""Benchmark of the MPC planning hot path (BASELINE.json metric: dynamics-model
candidate-steps/sec; CEM plan latency p50).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--engine auto|fp32|fp16|bf16]
                    [--workload cheetah|walker|cartpole|humanoid] [--no-extras]

One "step" = one whole CEM plan (sample -> rollout+cost -> top-k -> refit, I iterations, then
the chosen plan is emitted) on synthetic inputs of BASELINE config 3: cheetah-run shape
(obs 17, act 6), dynamics MLP hidden 200 with random-init weights, N=16384 candidates per GPU,
H=30, I=5, k=10%.  With --gpus N>1 (launched under torchrun, one rank per GPU, NCCL) the
population is sharded: N_total = 16384*N candidates, one elite exchange per iteration
(weak scaling: per-GPU work fixed).

Prints ONE JSON line (rank 0).  `value` is device-resident whole-job throughput (CUDA events,
max over ranks); `e2e` is the same metric through the host-buffer C-ABI call (`mbrl_plan`:
pinned H2D of s0 and D2H of the plan inside the timed region); `roofline` is the rollout
kernel against the measured tensor peak; `hbm_kernels` are the HBM-bound kernels (sampler,
top-k, refit) against the measured copy bandwidth; `cpu_baseline` is the oracle port of the
reference's CPU planner timed on this box's host cores.  Outside the headline timing the line
also carries, as extra keys, the other BASELINE configs at this GPU count:
`cfg4_strong` (walker-walk N=131072 split over the N GPUs, population-sharded), `cfg5_env`
(humanoid-run 1024 environments x 2048 candidates, hidden 512, split over the N GPUs with no
collective) and, for N>1, `sharded_equals_unsharded` (every rank holds the same plan and rank
0's equals an unsharded plan over the whole population, bit for bit).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "dynamics-model candidate-steps/sec (CEM plan, cheetah-run shape)"
UNIT = "candidate-steps/s"
WORKLOADS = {
    # BASELINE.json configs[2]: the configuration the metric is quoted on (fits one GPU)
    "cheetah": dict(name="cheetah-run CEM N=16384 H=30 I=5 hidden=200 (BASELINE configs[2])",
                    O=17, A=6, U=200, N=16384, H=30, I=5, E=1, elite_frac=0.1),
    # BASELINE.json configs[3] per-GPU shard at 8 GPUs (131072 / 8)
    "walker": dict(name="walker-walk CEM N=16384/GPU H=30 I=5 hidden=200 (BASELINE configs[3] shard)",
                   O=24, A=6, U=200, N=16384, H=30, I=5, E=1, elite_frac=0.1),
    "cartpole": dict(name="cartpole-swingup CEM N=4096 H=30 I=5 hidden=50 (BASELINE configs[1])",
                     O=5, A=1, U=50, N=4096, H=30, I=5, E=1, elite_frac=0.1),
    # BASELINE.json configs[4], one GPU's shard at 8 GPUs (1024 / 8 environments)
    "humanoid": dict(name="humanoid-run batched MPC 128 envs x N=2048 H=50 I=5 hidden=512 (BASELINE configs[4] shard)",
                     O=67, A=21, U=512, N=2048, H=50, I=5, E=128, elite_frac=0.1),
}
CFG4 = dict(O=24, A=6, U=200, N_total=131072, H=30, I=5, elite_frac=0.1)
CFG5 = dict(O=67, A=21, U=512, N=2048, H=50, I=5, E_total=1024, elite_frac=0.1)


def flops_per_cand_step(w):
    D = w["O"] + w["A"]
    return 2 * (D * w["U"] + w["U"] * w["U"] + w["U"] * w["O"])


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(bf16_tflops=d["bf16_tflops"], hbm_gbs=d["hbm_gbs"], source="measured")
    return dict(bf16_tflops=1590.0, hbm_gbs=6650.0, source="fallback")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        busy = sorted(sm)[len(sm) // 2:] if sm else []  # upper half ~ samples under load
        return dict(sm_mhz=statistics.median(busy) if busy else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference planner (bench.py's one permitted use of oracle/)
# ------------------------------------------------------------------------------------------
def cpu_reference_plan_time(w, iterations, reps, warm):
    """Times the reference-composed CEM: per iteration the reference-style
    _generate_trajectories restatement (autograd on, one sampler call, per-candidate Python list
    of views -- planners.py:189-216) with a Gaussian sampler closure, then stable argsort top-k
    and mean/std refit.  Returns (seconds per `iterations`-iteration plan, threads)."""
    import torch
    from oracle import planner_oracle as po

    p = po.synthetic_params(w["O"], w["A"], w["U"])
    for t in (p.W1, p.b1, p.W2, p.b2, p.W3, p.b3):
        t.requires_grad_(True)  # reference parameters are nn.Parameters; autograd stays on
    model, cost = po.params_as_callables(p)
    N, H, A = w["N"], w["H"], w["A"]
    k = max(1, int(w["elite_frac"] * N))
    times = []
    for rep in range(warm + reps):
        s0 = po.synthetic_state(p, rep)
        g = torch.Generator().manual_seed(rep)
        t0 = time.perf_counter()
        mu, sd = torch.zeros(H, A), torch.ones(H, A)
        for it in range(iterations):
            def gauss(batch_size, mu=mu, sd=sd):
                z = torch.randn(batch_size, A, generator=g)
                return torch.clamp(mu.repeat_interleave(N, 0) + sd.repeat_interleave(N, 0) * z, -1.0, 1.0)
            trajs, costs = po.reference_style_generate(s0, model, cost, gauss, H, N)
            elite = po.topk_stable(costs, k)
            acts = torch.stack([trajs[i][1] for i in elite], dim=1)
            mu, sd = acts.mean(1), acts.std(1, unbiased=False)
        dt = time.perf_counter() - t0
        if rep >= warm:
            times.append(dt)
    return statistics.median(times), torch.get_num_threads()


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the
    reference is pure Python and cannot travel to the GPU box) on this box's host cores."""
    if rank != 0:
        return
    import torch
    try:  # torchrun exports OMP_NUM_THREADS=1; the CPU arm is entitled to every host thread it can use
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, RuntimeError):
        pass
    w = dict(WORKLOADS[args.workload])
    # Bounded sample: every step is one full I-iteration plan over ONE GPU's share of the population
    # (N candidates), whatever --gpus is -- the CPU planner's candidate-steps/s does not depend on how
    # many shares there are, and a whole 8-share plan (13 s) times K steps would not end in minutes.
    iters = w["I"]
    if w.get("E", 1) > 1:  # batched environments: 8 environments' worth of rows, one iteration (hidden 512 is 7x the flops)
        w["N"], iters = w["N"] * 8, 1
    sec, threads = cpu_reference_plan_time(w, iters, args.steps, args.warmup)
    value = w["N"] * w["H"] * iters / sec
    sample = (f"each step = one reference-composed CEM plan, I={iters} iterations, over one GPU's share of the population "
              f"(N={w['N']} candidate rows, H={w['H']}) -- oracle port of planners.py:189-216, torch CPU fp32, autograd on, "
              f"per-candidate list build included; throughput per row does not depend on the number of shares")
    line = dict(
        impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
        ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload=w["name"] + (f" x{args.gpus} GPUs population-sharded, N_total={WORKLOADS[args.workload]['N'] * args.gpus}"
                                          if args.gpus > 1 else ""), l2="n/a (CPU)"),
        cpu_baseline=dict(value=value, unit=UNIT, cores=threads, kind="port", sample=sample),
        e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        gpu_launches=0,
    )
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def python_api_latency(prob, w, engine, k, states0, warmup, steps):
    """Plan latency through the reference-facing Python API: CEMPlanner.plan called the way
    MPCPolicy.get_action calls it (src/mbrl/agents.py:48-55) with callables wired like GoalStateAgent
    (agents.py:225-233) -- adaptor introspection, fingerprint check, ctypes call, host tensors back.
    p50 over `steps` (>= 100) consecutive calls with a fresh s0 each."""
    import torch
    from functools import partial
    from mbrl_b200 import CEMPlanner, planners

    class Net(torch.nn.Module):  # shaped like src/mbrl/models.py:96-104
        def __init__(self):
            super().__init__()
            self.linear1 = torch.nn.Linear(w["O"] + w["A"], w["U"])
            self.linear2 = torch.nn.Linear(w["U"], w["U"])
            self.linear3 = torch.nn.Linear(w["U"], w["O"])
            self.noise = None
    net = Net()
    with torch.no_grad():
        for lin, (W, b) in zip((net.linear1, net.linear2, net.linear3),
                               ((prob.W1, prob.b1), (prob.W2, prob.b2), (prob.W3, prob.b3))):
            lin.weight.copy_(W); lin.bias.copy_(b)
    stats = {"observations": {"mean": prob.mu_s, "std": prob.sd_s}, "actions": {"mean": prob.mu_a, "std": prob.sd_a}}

    def field(x, field_name, stats):
        raise RuntimeError("host callables are never invoked by the GPU planner")

    class StateCost:
        weights, goal_state, alpha = prob.cost_w, prob.goal, prob.alpha

    class ActionCost:
        alpha = prob.beta
    model = partial(net, normalize_state=partial(field, field_name="observations", stats=stats),
                    normalize_action=partial(field, field_name="actions", stats=stats),
                    unnormalize_state=partial(field, field_name="observations", stats=stats))
    cost = partial(field, state_cost=StateCost, action_cost=ActionCost)
    lat = []
    for i in range(warmup + steps):
        s0 = states0[i % len(states0)]
        t0 = time.perf_counter()
        _, actions = CEMPlanner.plan(s0, model, cost, None, w["H"], None, num_trajectories=w["N"],
                                     num_iterations=w["I"], num_elites=k, engine=engine, seed=i, return_states=False)
        first_action = actions[0].flatten()
        dt = time.perf_counter() - t0
        if i >= warmup:
            lat.append(dt)
    planners.clear_handles()
    assert first_action.shape == (w["A"],)
    return statistics.median(lat) * 1e3, len(lat)


def pick_engine(native, w, requested):
    if requested != "auto":
        return requested
    for eng in ("fp16", "bf16"):
        try:
            native.NativePlanner(w["O"], w["A"], w["U"], w["H"], 256, engine=eng).close()
            return eng
        except native.MbrlError:
            continue
    return "fp32"


def _pct(xs, q):
    xs = sorted(xs)
    return xs[min(len(xs) - 1, int(q * len(xs)))]


class Dist:
    """The torch.distributed plumbing bench.py needs (no-ops on one GPU)."""

    def __init__(self, world, dev):
        self.world, self.dev = world, dev

    def barrier(self):
        import torch
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_(self, values):
        """element-wise MAX over ranks of a list of floats"""
        import torch
        t = torch.tensor(values, dtype=torch.float64, device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()


def attach_shards(h, native, transport, rank, world, local_rank, dev):
    """Make handle h one shard of a population split over the ranks: NVLink peer memory when every GPU
    pair is P2P-accessible and every rank could open its peers' buffers, else the in-library ncclAllGather.
    Collective: every rank takes the same transport.  Returns the transport used."""
    import torch
    import torch.distributed as dist
    if transport == "p2p":
        ok = all(torch.cuda.can_device_access_peer(local_rank, r) for r in range(world) if r != local_rank)
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            transport = "nccl"
    if transport == "p2p" and not h.p2p_init(rank, world):
        transport = "nccl"  # some rank could not open a peer buffer: every rank detached, all take NCCL
    if transport == "nccl":
        h.comm_init(rank, world)
    return transport


def time_plans(plan, steps, warmup, flush, D):
    """W warm-up plans, then K plans each between a CUDA-event pair on the launching stream (L2 flushed
    before each, outside the pair), barrier + synchronize on both sides.  Returns per-plan ms, MAX over ranks."""
    import torch
    for i in range(warmup):
        plan(i)
    D.barrier()
    evs = []
    D.barrier()
    for i in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan(warmup + i)
        e1.record()
        evs.append((e0, e1))
    D.barrier()
    return D.max_([a.elapsed_time(b) for a, b in evs])


def hbm_kernel_lines(native, h, w, k, d_s0, dev, flush, hbm_gbs):
    """The HBM-bound kernels of a plan timed alone (CUDA events, L2 flushed): achieved GB/s over their
    ALGORITHMIC bytes against the measured copy bandwidth.  At these sizes (tens of KB to 12 MB) they are
    launch/latency-bound, which is why the production plan fuses the sampler into the rollout kernel and
    regenerates elite actions from Philox counters in the refit instead of reading them."""
    import torch
    O, A, N, H, I = w["O"], w["A"], w["N"], w["H"], w["I"]
    mu = torch.zeros(1, H, A, device=dev)
    sd = torch.ones(1, H, A, device=dev)
    costs = h.rollout(d_s0, native.SAMPLE_GAUSSIAN, 1, 0, d_mu=mu, d_sd=sd)[0]
    idx, _, _ = native.topk(costs, k, 1)

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts) * 1e3  # us

    out = []
    for name, fn, nbytes, note in (
        ("sample_kernel (materialising Philox sampler + clip)", lambda: h.sample(native.SAMPLE_GAUSSIAN, 1, 0, d_mu=mu, d_sd=sd),
         4 * H * N * A, "writes 4*H*N*A bytes; the production plan never materialises actions (sampler fused into the rollout: 0 B)"),
        ("topk_select_kernel (elite select + best-ever)", lambda: native.topk(costs, k, 1),
         4 * N + 8 * k, "reads 4*N bytes of costs, writes 4*k indices + 4*k costs"),
        ("refit_kernel (mean/std over elites)", lambda: h.refit(idx, k, native.SAMPLE_GAUSSIAN, 1, 0, d_mu=mu, d_sd=sd),
         4 * k * H * A + 8 * H * A, "algorithmic bytes of the gather formulation (4*k*H*A elite actions in, 8*H*A out); "
                                   "the kernel regenerates the elites' actions from Philox counters and reads 4*k index bytes instead"),
    ):
        us = timed(fn)
        gbs = nbytes / (us * 1e-6) / 1e9
        out.append(dict(kernel=name, us=us, algorithmic_bytes=nbytes, achieved_gbs=gbs, peak_gbs=hbm_gbs,
                        frac=gbs / hbm_gbs, note=note))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--engine", default="auto", choices=["auto", "fp32", "fp16", "bf16"])
    ap.add_argument("--workload", default="cheetah", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg4_strong / cfg5_env / equality extras")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"],
                    help="elite exchange of the population-sharded loop: NVLink peer stores (CUDA IPC) or ncclAllGather")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from mbrl_b200 import native
    from mbrl_b200.sharding import env_shard
    from mbrl_b200.synthetic import synthetic_problem, synthetic_state

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    D = Dist(world, dev)

    w = WORKLOADS[args.workload]
    O, A, U, N, H, I, E = w["O"], w["A"], w["U"], w["N"], w["H"], w["I"], w["E"]
    env_mode = E > 1  # batched independent environments: sharded by environment, no collective
    n_total = N * (1 if env_mode else world)
    k = max(1, int(w["elite_frac"] * n_total))
    engine = pick_engine(native, w, args.engine)
    peaks = measured_peaks()

    # synthetic problem; the CPU arm's oracle generates the identical values
    prob = p = synthetic_problem(O, A, U)
    h = native.NativePlanner(O, A, U, H, N, E, I, k, engine, local_rank)
    h.load_problem(prob)
    transport = None
    if world > 1 and not env_mode:
        # the sharded CEM loop runs on the stream inside the library; elite exchange over NVLink peer
        # memory (default) or an in-library ncclAllGather
        transport = attach_shards(h, native, args.transport, rank, world, local_rank, dev)
    n_states = args.warmup + args.steps
    states0 = torch.stack([torch.stack([synthetic_state(p, c * E + e + (rank * 100003 if env_mode else 0)) for e in range(E)])
                           for c in range(n_states)]).float()  # [plans, E, O]
    d_states0 = states0.to(dev)
    d_out_s = torch.empty(E, H, O, device=dev)
    d_out_a = torch.empty(E, H, A, device=dev)
    d_info = torch.zeros(E, 4, dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def plan_resident(i):
        h.plan_device(d_states0[i % n_states], d_out_s, d_out_a, d_info, iterations=I, elites=k,
                      mode=native.SAMPLE_GAUSSIAN, seed=i, env_offset=(rank * E if env_mode else 0))

    # ---- device-resident throughput ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # runs through warm-up, the timed regions and a short load tail (see below)
    step_ms = time_plans(plan_resident, args.steps, args.warmup, flush, D)
    total_ms = sum(step_ms)
    cand_steps_per_plan = n_total * E * (world if env_mode else 1) * H * I
    value = cand_steps_per_plan * args.steps / (total_ms * 1e-3)

    # ---- end to end through the host-buffer C-ABI call ----
    e2e_lat, e2e_lat_actions = [], []
    for actions_only, sink in ((False, e2e_lat), (True, e2e_lat_actions)):
        for i in range(args.warmup + args.steps):
            s0 = states0[i].numpy()
            if world > 1:
                D.barrier()  # ranks enter every timed call together (a sharded plan waits for its slowest rank anyway)
            t0 = time.perf_counter()
            h.plan(s0, iterations=I, elites=k, mode=native.SAMPLE_GAUSSIAN, seed=i, actions_only=actions_only,
                   env_offset=(rank * E if env_mode else 0))
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                sink.append(dt)
    e2e_ms = D.max_([x * 1e3 for x in e2e_lat])
    e2e_value = cand_steps_per_plan * args.steps / (sum(e2e_ms) * 1e-3)
    py_api = None
    if world == 1 and not env_mode:
        py_api = python_api_latency(prob, w, engine, k, states0[:, 0], args.warmup, max(args.steps, 100))

    # A timed region of K sub-millisecond plans is shorter than nvidia-smi's sampling period: keep the
    # same load running (untimed) until the sampler has a few readings under load.  The number of tail
    # plans derives from the all-reduced step time, so every rank runs the SAME count (a sharded plan
    # contains an elite exchange: a rank that planned more often than its peers would wait forever).
    n_tail = min(20000, int(1500.0 / max(total_ms / args.steps, 1e-3)) + 1)
    for j in range(n_tail):
        plan_resident(args.warmup + (j % args.steps))
        if (j + 1) % 50 == 0:
            torch.cuda.synchronize()
    D.barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---- dominant kernel alone: rollout + cost (tensor-bound) ----
    mu = torch.zeros(E, H, A, device=dev)
    sd = torch.ones(E, H, A, device=dev)
    for _ in range(3):
        h.rollout(d_states0[0], native.SAMPLE_GAUSSIAN, 1, 0, d_mu=mu, d_sd=sd)
    torch.cuda.synchronize()
    kev = []
    for i in range(20 if not env_mode else 5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h.rollout(d_states0[0], native.SAMPLE_GAUSSIAN, 1, i % I, d_mu=mu, d_sd=sd)
        e1.record()
        kev.append((e0, e1))
    torch.cuda.synchronize()
    k_ms = statistics.mean(a.elapsed_time(b) for a, b in kev)
    # the same kernel the way a plan runs it: I launches back to back after ONE L2 flush (its prologue -- barrier
    # init, TMEM allocation, the weight fetch -- overlaps the predecessor's tail under programmatic dependent launch)
    kev2 = []
    for rep in range(6 if not env_mode else 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(I):
            h.rollout(d_states0[0], native.SAMPLE_GAUSSIAN, 1, i, d_mu=mu, d_sd=sd)
        e1.record()
        kev2.append((e0, e1))
    torch.cuda.synchronize()
    k_ms_b2b = statistics.mean(a.elapsed_time(b) for a, b in kev2) / I
    alg_flops = N * E * H * flops_per_cand_step(w)
    achieved = alg_flops / (k_ms * 1e-3) / 1e12
    traffic, traffic_source = None, None
    tpath = os.path.join(ROOT, "profiles", "rollout_traffic.json")  # DRAM bytes per launch from ncu --set full captures
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        entry = tj.get(args.workload, {}).get(engine)
        if entry:
            traffic = entry["dram_bytes_per_launch"]
            traffic_source = entry["source"] + " (static: read from profiles/rollout_traffic.json, not re-measured in this run)"
    roofline = dict(bound="tensor", kernel="rollout+cost (%s engine)" % engine, achieved=achieved,
                    peak=peaks["bf16_tflops"], unit="TFLOP/s", frac=achieved / peaks["bf16_tflops"], traffic=traffic,
                    traffic_source=traffic_source, peak_source=peaks["source"] + " cuBLAS bf16 burst", kernel_ms=k_ms,
                    kernel_ms_how="mean of single launches, each between its own CUDA-event pair after an L2 flush (includes the "
                                  "launch latency and the un-overlapped prologue: the conservative figure, same method as round 1)",
                    kernel_ms_back_to_back=k_ms_b2b,
                    frac_back_to_back=alg_flops / (k_ms_b2b * 1e-3) / 1e12 / peaks["bf16_tflops"],
                    back_to_back_how="%d launches back to back between one event pair after one L2 flush, as inside a plan" % I,
                    algorithmic_flops_per_launch=alg_flops)

    hbm_kernels = None
    if world == 1 and not env_mode:
        hbm_kernels = hbm_kernel_lines(native, h, w, k, d_states0[0], dev, flush, peaks["hbm_gbs"])

    # ---------------------------------------------------------------------------------------
    # extras, outside the headline timing
    # ---------------------------------------------------------------------------------------
    equal = None
    if world > 1 and not env_mode and not args.no_extras:
        # every rank holds the same plan, and rank 0's equals an UNSHARDED plan over the whole population
        s0 = states0[0].numpy()
        out = h.plan(s0, I, k, native.SAMPLE_GAUSSIAN, seed=12345, want_dist=True)
        sig = np.concatenate([out["actions"].ravel(), out["states"].ravel(), out["mu"].ravel(), out["sd"].ravel(),
                              out["info"]["best_cost"], out["info"]["best_index"].astype(np.float32),
                              out["info"]["best_iteration"].astype(np.float32)]).astype(np.float32)
        mine = torch.from_numpy(sig).to(dev)
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        same = torch.tensor([1 if torch.equal(mine.view(torch.int32), ref.view(torch.int32)) else 0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        equal = dict(ranks_identical=bool(same.item()))
        if rank == 0:
            full = native.NativePlanner(O, A, U, H, n_total, 1, I, k, engine, local_rank)
            full.load_problem(prob)
            full.set_refit_segments(world)  # same summation order as the sharded refit (per-rank partials, rank order)
            want = full.plan(s0, I, k, native.SAMPLE_GAUSSIAN, seed=12345, want_dist=True)
            eq = all(np.array_equal(out[key].view(np.int32), want[key].view(np.int32)) for key in ("actions", "states", "mu", "sd"))
            eq = eq and all(np.array_equal(out["info"][key], want["info"][key]) for key in ("best_cost", "best_index", "best_iteration"))
            equal["rank0_equals_unsharded_plan"] = bool(eq)
            equal["n_total"] = n_total
            full.close()
        D.barrier()

    gd = None
    if world == 1 and not env_mode and not args.no_extras:
        # the batched GradientDescentPlanner (planners.py:28-137), 40 Adam iterations with
        # back-propagation through the H-step rollout, for 1 and for 128 restarts; CPU: the oracle port, 1 restart
        rng = np.random.default_rng(0)
        init = rng.uniform(-1, 1, (128, H, A)).astype(np.float32)
        s0g = states0[0, 0].numpy()
        for B in (1, 128):
            h.plan_gd(s0g, init[:B], iterations=40, stop_condition=0.0)
        tg = {}
        for B in (1, 128):
            ts = []
            for _ in range(5):
                t0 = time.perf_counter(); h.plan_gd(s0g, init[:B], iterations=40, stop_condition=0.0); ts.append(time.perf_counter() - t0)
            tg[B] = statistics.median(ts) * 1e3
        gd = dict(workload=f"GradientDescentPlanner, {w['name'].split(' CEM')[0]} shape, H={H}, 40 Adam iterations (no early stop)",
                  ms_per_plan_1_restart=tg[1], ms_per_plan_128_restarts=tg[128], api="mbrl_plan_gd (host buffers)")
        if not args.no_cpu_baseline:
            from oracle import planner_oracle as po
            pp = po.synthetic_params(O, A, U)
            t0 = time.perf_counter()
            po.gd_plan(pp, po.synthetic_state(pp, 0), torch.from_numpy(init[0]), H, 40, 0.0)
            gd["cpu_oracle_ms_per_plan_1_restart"] = (time.perf_counter() - t0) * 1e3

    cfg4 = cfg5 = None
    if not args.no_extras and not env_mode:
        D.barrier()
        h.close()
        del flush
        torch.cuda.empty_cache()
        flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
        ksteps, kwarm = min(args.steps, 10), 3
        # ---- BASELINE configs[3]: walker-walk N=131072 split over the GPUs (strong scaling) ----
        c = CFG4
        if c["N_total"] % world == 0:
            n_l = c["N_total"] // world
            k4 = int(c["elite_frac"] * c["N_total"])
            prob4 = synthetic_problem(c["O"], c["A"], c["U"])
            h4 = native.NativePlanner(c["O"], c["A"], c["U"], c["H"], n_l, 1, c["I"], k4, engine, local_rank)
            h4.load_problem(prob4)
            tr4 = attach_shards(h4, native, args.transport, rank, world, local_rank, dev) if world > 1 else None
            s4 = torch.stack([synthetic_state(prob4, i) for i in range(kwarm + ksteps)]).float().to(dev)
            o_s, o_a = torch.empty(1, c["H"], c["O"], device=dev), torch.empty(1, c["H"], c["A"], device=dev)
            inf4 = torch.zeros(1, 4, dtype=torch.int32, device=dev)
            ms4 = time_plans(lambda i: h4.plan_device(s4[i:i + 1], o_s, o_a, inf4, iterations=c["I"], elites=k4,
                                                      mode=native.SAMPLE_GAUSSIAN, seed=i), ksteps, kwarm, flush, D)
            cs4 = c["N_total"] * c["H"] * c["I"]
            cfg4 = dict(workload=f"walker-walk CEM N_total={c['N_total']} ({n_l}/GPU x {world}) H={c['H']} I={c['I']} hidden={c['U']} (BASELINE configs[3])",
                        scaling="strong", n_gpus=world, ms_per_plan=statistics.mean(ms4), ms_per_plan_p50=statistics.median(ms4),
                        value=cs4 * ksteps / (sum(ms4) * 1e-3), unit=UNIT, steps=ksteps, warmup=kwarm, transport=tr4,
                        algorithmic_tflops=cs4 * flops_per_cand_step(c) / (statistics.mean(ms4) * 1e-3) / 1e12)
            h4.close()
        # ---- BASELINE configs[4]: humanoid-run 1024 envs x 2048 candidates, hidden 512, env-sharded ----
        c = CFG5
        first, e_l = env_shard(c["E_total"], rank, world)
        k5 = int(c["elite_frac"] * c["N"])
        prob5 = synthetic_problem(c["O"], c["A"], c["U"])
        try:
            eng5 = engine if engine == "fp32" else pick_engine(native, dict(O=c["O"], A=c["A"], U=c["U"], H=c["H"]), "auto")
            h5 = native.NativePlanner(c["O"], c["A"], c["U"], c["H"], c["N"], e_l, c["I"], k5, eng5, local_rank)
            h5.load_problem(prob5)
            k5steps = min(args.steps, 3)
            s5 = torch.stack([torch.stack([synthetic_state(prob5, (first + e) * 7 + i) for e in range(e_l)])
                              for i in range(2 + k5steps)]).float().to(dev)
            o_s, o_a = torch.empty(e_l, c["H"], c["O"], device=dev), torch.empty(e_l, c["H"], c["A"], device=dev)
            inf5 = torch.zeros(e_l, 4, dtype=torch.int32, device=dev)
            ms5 = time_plans(lambda i: h5.plan_device(s5[i], o_s, o_a, inf5, iterations=c["I"], elites=k5,
                                                      mode=native.SAMPLE_GAUSSIAN, seed=i, env_offset=first), k5steps, 2, flush, D)
            cs5 = c["E_total"] * c["N"] * c["H"] * c["I"]
            cfg5 = dict(workload=f"humanoid-run batched MPC {c['E_total']} envs ({e_l}/GPU x {world}) x N={c['N']} H={c['H']} I={c['I']} hidden={c['U']} (BASELINE configs[4])",
                        scaling="strong", n_gpus=world, engine=eng5, collective="none (environment sharding)",
                        ms_per_plan=statistics.mean(ms5), ms_per_iteration=statistics.mean(ms5) / c["I"],
                        value=cs5 * k5steps / (sum(ms5) * 1e-3), unit=UNIT, steps=k5steps, warmup=2,
                        algorithmic_tflops=cs5 * flops_per_cand_step(c) / (statistics.mean(ms5) * 1e-3) / 1e12,
                        frac_of_tensor_peak_per_gpu=cs5 * flops_per_cand_step(c) / (statistics.mean(ms5) * 1e-3) / 1e12 / world / peaks["bf16_tflops"])
            h5.close()
        except native.MbrlError as exc:
            cfg5 = dict(error=str(exc))

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        wc = dict(w); wc["N"] = N if not env_mode else N * 8  # batched environments: 8 environments' worth of rows
        reps = 3 if not env_mode else 1
        sec, threads = cpu_reference_plan_time(wc, I if not env_mode else 1, reps=reps, warm=1 if not env_mode else 0)
        it_cpu = I if not env_mode else 1
        cpu = dict(value=wc["N"] * H * it_cpu / sec, unit=UNIT, cores=threads, kind="port",
                   sample=f"{reps} timed reference-composed CEM plan(s) (I={it_cpu}, {wc['N']} candidate rows, H={H}), "
                          f"oracle port of planners.py:189-216 incl. autograd + Python list build; {sec * 1e3:.0f} ms/plan",
                   ms_per_plan=sec * 1e3)

    # init + I x (rollout, select) + (I-1) refits + replay; population-sharded: init + I x (rollout, local select,
    # merge select) + (I-1) distributed refits + replay + the flag kernel
    launches_per_plan = (1 + 2 * I + (I - 1) + 1) if (world == 1 or env_mode) else (1 + 3 * I + (I - 1) + 2)
    cfg = dict(workload=w["name"] + (f" x{world} GPUs population-sharded, N_total={n_total}" if (world > 1 and not env_mode) else "")
               + (f" x{world} GPUs environment-sharded" if (world > 1 and env_mode) else ""),
               engine=engine, elites=k, l2="flushed between timed plans (256 MiB write)",
               parallelism=("single GPU" if world == 1 else
                            ("environment-sharded x%d, no collective" % world if env_mode else
                             "population-sharded x%d, one (cost, index) elite exchange per iteration over %s"
                             % (world, "NVLink peer memory" if transport == "p2p" else "ncclAllGather"))))
    line = dict(
        metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype={"fp32": "f32", "fp16": "f16 operands / f32 accumulate", "bf16": "bf16 operands / f32 accumulate"}[engine],
        data="synthetic", config=cfg,
        plan_latency_ms_p50=statistics.median(step_ms), plan_latency_ms_mean=statistics.mean(step_ms),
        plan_latency_ms_p99=_pct(step_ms, 0.99), plan_latency_ms_min=min(step_ms),
        clocks=clocks,
        e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=4 * O * E, d2h_bytes_per_step=E * (4 * H * (O + A) + 16),
                 latency_ms_p50=statistics.median(e2e_ms), latency_ms_mean=statistics.mean(e2e_ms),
                 latency_ms_p50_actions_only=(statistics.median(e2e_lat_actions) * 1e3 if e2e_lat_actions else None),
                 python_api_first_action_latency_ms_p50=(py_api[0] if py_api else None),
                 python_api_calls=(py_api[1] if py_api else None),
                 api="mbrl_plan (host buffers)" + ("" if world == 1 or env_mode else ", population-sharded (in-library elite exchange: %s)" % transport),
                 transfer="host buffers -> the handle's pinned, device-mapped staging buffers; the plan's first kernel reads s0 from "
                          "there and its last kernels store the plan there (PCIe reads / posted writes inside the timed region; "
                          "copy-engine transfers above 256 KB)"),
        gpu_launches=launches_per_plan * args.steps,
        roofline=roofline, hbm_kernels=hbm_kernels, cpu_baseline=cpu,
        sharded_equals_unsharded=(None if equal is None else bool(equal.get("ranks_identical") and equal.get("rank0_equals_unsharded_plan"))),
        sharded_check=equal, cfg4_strong=cfg4, cfg5_env=cfg5, gradient_planner=gd,
    )
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
