#!/usr/bin/env python
"""Benchmark of the MPC planning hot path (BASELINE.json metric: dynamics-model
candidate-steps/sec; CEM plan latency p50).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--engine auto|fp32|fp16|bf16]

One "step" = one whole CEM plan (sample -> rollout+cost -> top-k -> refit, I iterations, then
the chosen plan is emitted) on synthetic inputs of BASELINE config 3: cheetah-run shape
(obs 17, act 6), dynamics MLP hidden 200 with random-init weights, N=16384 candidates per GPU,
H=30, I=5, k=10%.  With --gpus N>1 (launched under torchrun, one rank per GPU, NCCL) the
population is sharded: N_total = 16384*N candidates, one elite all-gather per iteration
(weak scaling: per-GPU work fixed).

Prints ONE JSON line (rank 0).  `value` is device-resident whole-job throughput (CUDA events,
max over ranks); `e2e` is the same metric through the host-buffer C-ABI call (`mbrl_plan`:
pinned H2D of s0 and D2H of the plan inside the timed region); `roofline` is the rollout
kernel against the measured tensor peak; `cpu_baseline` is the oracle port of the reference's
CPU planner timed on this box's host cores.
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
                    O=17, A=6, U=200, N=16384, H=30, I=5, elite_frac=0.1),
    # BASELINE.json configs[3] per-GPU shard at 8 GPUs (131072 / 8)
    "walker": dict(name="walker-walk CEM N=16384/GPU H=30 I=5 hidden=200 (BASELINE configs[3] shard)",
                   O=24, A=6, U=200, N=16384, H=30, I=5, elite_frac=0.1),
    "cartpole": dict(name="cartpole-swingup CEM N=4096 H=30 I=5 hidden=50 (BASELINE configs[1])",
                     O=5, A=1, U=50, N=4096, H=30, I=5, elite_frac=0.1),
}


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
    import numpy as np
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
    w["N"] = w["N"] * max(1, args.gpus)  # same whole-job population as the B200 arm at --gpus N
    iters = w["I"] if (args.steps + args.warmup) <= 40 and args.gpus == 1 else 1
    sec, threads = cpu_reference_plan_time(w, iters, args.steps, args.warmup)
    value = w["N"] * w["H"] * iters / sec
    sample = (f"each step = one reference-composed CEM plan restricted to {iters} iteration(s) of "
              f"N={w['N']} H={w['H']} (oracle port of planners.py:189-216, torch CPU fp32, autograd on, list build included)")
    line = dict(
        impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
        ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload=w["name"] + (f" x{args.gpus} (N_total={w['N']})" if args.gpus > 1 else ""), l2="n/a (CPU)"),
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
    (agents.py:225-233) -- adaptor introspection, fingerprint check, ctypes call, host tensors back."""
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
        t0 = time.perf_counter()
        _, actions = CEMPlanner.plan(states0[i], model, cost, None, w["H"], None, num_trajectories=w["N"],
                                     num_iterations=w["I"], num_elites=k, engine=engine, seed=i, return_states=False)
        first_action = actions[0].flatten()
        dt = time.perf_counter() - t0
        if i >= warmup:
            lat.append(dt)
    planners.clear_handles()
    assert first_action.shape == (w["A"],)
    return statistics.median(lat) * 1e3


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--engine", default="auto", choices=["auto", "fp32", "fp16", "bf16"])
    ap.add_argument("--workload", default="cheetah", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    from mbrl_b200 import PlanningProblem, native
    from mbrl_b200.sharding import NativeOps, PopulationShardedCEM

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    w = WORKLOADS[args.workload]
    O, A, U, N, H, I = w["O"], w["A"], w["U"], w["N"], w["H"], w["I"]
    n_total = N * world
    k = max(1, int(w["elite_frac"] * n_total))
    engine = pick_engine(native, w, args.engine)

    # synthetic problem; the CPU arm's oracle generates the identical values
    from mbrl_b200.synthetic import synthetic_problem, synthetic_state
    prob = p = synthetic_problem(O, A, U)
    h = native.NativePlanner(O, A, U, H, N, 1, I, k, engine, local_rank)
    h.load_problem(prob)
    if world > 1:
        # the sharded CEM loop runs on the stream inside the library; elite exchange over NVLink peer
        # memory (default) or an in-library ncclAllGather
        transport = args.transport
        if transport == "p2p":
            # peer-memory exchange needs every GPU pair of this node to be P2P-accessible (rank r drives
            # device r under torchrun); agree on it collectively, else every rank takes the NCCL transport
            ok = all(torch.cuda.can_device_access_peer(local_rank, r) for r in range(world) if r != local_rank)
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                transport = "nccl"
        if transport == "p2p" and not h.p2p_init(rank, world):
            transport = "nccl"  # some rank could not open a peer buffer: every rank detached, all take NCCL
        if transport == "nccl":
            h.comm_init(rank, world)
        args.transport = transport
    states0 = torch.stack([synthetic_state(p, c) for c in range(args.warmup + args.steps)]).float()
    d_states0 = states0.to(dev)
    d_out_s = torch.empty(1, H, O, device=dev)
    d_out_a = torch.empty(1, H, A, device=dev)
    d_info = torch.zeros(1, 4, dtype=torch.int32, device=dev)
    sharded = None  # (the Python host loop mbrl_b200.sharding.PopulationShardedCEM remains as the tested reference of this logic)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def plan_resident(i):
        if sharded is None:
            h.plan_device(d_states0[i:i + 1], d_out_s, d_out_a, d_info, iterations=I, elites=k,
                          mode=native.SAMPLE_GAUSSIAN, seed=i)
        else:
            sharded.plan(d_states0[i:i + 1], I, k, seed=i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # runs through warm-up, the timed regions and a short load tail (see below)
    for i in range(args.warmup):
        plan_resident(i)
    barrier()
    evs = []
    barrier()
    for i in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the event pair)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan_resident(args.warmup + i)
        e1.record()
        evs.append((e0, e1))
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    cand_steps_per_plan = n_total * H * I
    value = cand_steps_per_plan * args.steps / (total_ms * 1e-3)

    # ---- end to end through the host-buffer C-ABI call ----
    e2e_lat, e2e_lat_actions = [], []
    if sharded is None:
        for i in range(args.warmup + args.steps):
            s0 = states0[i].numpy()
            t0 = time.perf_counter()
            out = h.plan(s0, iterations=I, elites=k, mode=native.SAMPLE_GAUSSIAN, seed=i)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                e2e_lat.append(dt)
        for i in range(args.warmup + args.steps):  # informational: first-action latency without the state replay
            s0 = states0[i].numpy()
            t0 = time.perf_counter()
            out = h.plan(s0, iterations=I, elites=k, mode=native.SAMPLE_GAUSSIAN, seed=i, actions_only=True)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                e2e_lat_actions.append(dt)
    else:
        pin = states0.pin_memory()
        for i in range(args.warmup + args.steps):
            barrier()
            t0 = time.perf_counter()
            d = pin[i:i + 1].to(dev, non_blocking=True)
            res = sharded.plan(d, I, k, seed=i)
            _ = res["actions"].cpu(), res["states"].cpu()
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                e2e_lat.append(dt)
    py_api_ms = None
    if world == 1:
        py_api_ms = python_api_latency(prob, w, engine, k, states0, args.warmup, args.steps)
    e2e_total = torch.tensor([sum(e2e_lat)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_total, op=dist.ReduceOp.MAX)
    e2e_value = cand_steps_per_plan * args.steps / float(e2e_total.item())

    # A timed region of K sub-millisecond plans is shorter than nvidia-smi's sampling period: keep the
    # same load running (untimed) until the sampler has a few readings under load.
    # The number of tail plans is derived from the all-reduced step time, so every rank runs the SAME
    # count: a sharded plan contains an elite exchange, and a rank that planned more often than its
    # peers would wait for exchanges that never come.
    n_tail = min(20000, int(1500.0 / max(total_ms / args.steps, 1e-3)) + 1)
    for j in range(n_tail):
        plan_resident(args.warmup + (j % args.steps))
        if (j + 1) % 50 == 0:
            torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---- dominant kernel alone: rollout + cost (tensor-bound) ----
    mu = torch.zeros(1, H, A, device=dev)
    sd = torch.ones(1, H, A, device=dev)
    for _ in range(3):
        h.rollout(d_states0[:1], native.SAMPLE_GAUSSIAN, 1, 0, d_mu=mu, d_sd=sd)
    torch.cuda.synchronize()
    kev = []
    for i in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h.rollout(d_states0[:1], native.SAMPLE_GAUSSIAN, 1, i % I, d_mu=mu, d_sd=sd)
        e1.record()
        kev.append((e0, e1))
    torch.cuda.synchronize()
    k_ms = statistics.mean(a.elapsed_time(b) for a, b in kev)
    peaks = measured_peaks()
    alg_flops = N * H * flops_per_cand_step(w)
    achieved = alg_flops / (k_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "rollout_traffic.json")  # dram bytes per launch from the ncu --set full capture
    if os.path.exists(tpath) and args.workload == "cheetah":
        with open(tpath) as f:
            traffic = json.load(f).get(engine)
    roofline = dict(bound="tensor", kernel="rollout+cost (%s engine)" % engine, achieved=achieved,
                    peak=peaks["bf16_tflops"], unit="TFLOP/s", frac=achieved / peaks["bf16_tflops"], traffic=traffic,
                    peak_source=peaks["source"] + " cuBLAS bf16 burst", kernel_ms=k_ms,
                    algorithmic_flops_per_launch=alg_flops)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sec, threads = cpu_reference_plan_time(w, I, reps=3, warm=1)
        cpu = dict(value=N * H * I / sec, unit=UNIT, cores=threads, kind="port",
                   sample=f"3 timed reference-composed CEM plans (I={I}, N={N}, H={H}) after 1 warm-up, "
                          f"oracle port of planners.py:189-216 incl. autograd + Python list build; {sec * 1e3:.0f} ms/plan",
                   ms_per_plan=sec * 1e3)

    launches_per_plan = (1 + 2 * I + (I - 1) + 1) if world == 1 else (1 + 6 * I + (I - 1) + 1)
    line = dict(
        metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype={"fp32": "f32", "fp16": "f16 operands / f32 accumulate", "bf16": "bf16 operands / f32 accumulate"}[engine],
        data="synthetic",
        config=dict(workload=w["name"] + (f" x{world} GPUs population-sharded, N_total={n_total}" if world > 1 else ""),
                    engine=engine, elites=k, l2="flushed between timed plans (256 MiB write)",
                    parallelism=("population-sharded x%d, one (cost, index) elite exchange per iteration over %s" % (world, "NVLink peer memory" if args.transport == "p2p" else "ncclAllGather")) if world > 1 else "single GPU"),
        plan_latency_ms_p50=statistics.median(step_ms),
        clocks=clocks,
        e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=4 * O, d2h_bytes_per_step=4 * H * (O + A) + 16,
                 latency_ms_p50=statistics.median(e2e_lat) * 1e3,
                 latency_ms_p50_actions_only=(statistics.median(e2e_lat_actions) * 1e3 if e2e_lat_actions else None),
                 python_api_first_action_latency_ms_p50=py_api_ms,
                 api="mbrl_plan (host buffers)" + ("" if world == 1 else ", population-sharded (in-library elite exchange: %s)" % args.transport)),
        gpu_launches=launches_per_plan * args.steps,
        roofline=roofline,
        cpu_baseline=cpu,
    )
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
