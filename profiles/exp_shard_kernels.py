"""Single-GPU timing of the per-iteration kernels of the 8-way population-sharded loop (cfg 3 per
GPU: N=16384, global k=13107): local top-k, merge top-k over the gathered elites, chunked refit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mbrl_b200 import native
from mbrl_b200.synthetic import synthetic_problem

O, A, U, H, N = 17, 6, 200, 30, 16384
world, k = 8, 13107
kl = int(k / world + 8 * (k / world) ** 0.5 + 64)
prob = synthetic_problem(O, A, U)
h = native.NativePlanner(O, A, U, H, N, 1, 5, k, "fp16")
h.load_problem(prob)
g = torch.Generator(device="cuda").manual_seed(0)
costs = torch.rand(N, device="cuda", generator=g)
gathered = torch.rand(world * kl, device="cuda", generator=g)
mu = torch.zeros(1, H, A, device="cuda"); sd = torch.ones(1, H, A, device="cuda")
elite_big = torch.sort(torch.randperm(N * world, device="cuda", generator=g)[:k]).values.int().view(1, -1)
elite_small = torch.sort(torch.randperm(N, device="cuda", generator=g)[:1638]).values.int().view(1, -1)

def timed(fn, n=15, batch=40):
    """GPU time per launch: `batch` back-to-back launches captured in a CUDA graph (no host launch
    cost in the timed region; launch gaps between dependent kernels included), median of n replays."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(batch): fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / batch)
    ts.sort()
    return ts[len(ts) // 2]

print("local top-k   n=%d k=%d : %.1f us" % (N, kl, timed(lambda: native.topk(costs, kl, 1))))
print("local top-k   n=%d k=1638 : %.1f us" % (N, timed(lambda: native.topk(costs, 1638, 1))))
print("merge top-k   n=%d k=%d : %.1f us" % (world * kl, k, timed(lambda: native.topk(gathered, k, 1))))
print("refit k=%d (chunks=%d) : %.1f us" % (k, (k + 2047) // 2048, timed(lambda: h.refit(elite_big, k, native.SAMPLE_GAUSSIAN, 1, 0, d_mu=mu, d_sd=sd))))
print("refit k=1638 : %.1f us" % timed(lambda: h.refit(elite_small, 1638, native.SAMPLE_GAUSSIAN, 1, 0, d_mu=mu, d_sd=sd)))
