"""Rollout-kernel time by action source (Philox Gaussian in-kernel vs injected buffer): how much
the in-kernel sampler's ALU work interferes with the MMA/epilogue pipeline.  Run on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mbrl_b200 import native
from mbrl_b200.synthetic import synthetic_problem, synthetic_state

O, A, U, H, N = 17, 6, 200, 30, 16384
prob = synthetic_problem(O, A, U)
h = native.NativePlanner(O, A, U, H, N, engine=sys.argv[1] if len(sys.argv) > 1 else "fp16")
h.load_problem(prob)
s0 = synthetic_state(prob, 0)[None].cuda()
mu = torch.zeros(1, H, A, device="cuda"); sd = torch.ones(1, H, A, device="cuda")
inj = h.sample(native.SAMPLE_GAUSSIAN, 1, 0, mu, sd)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timed(fn, n=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

print("gaussian (in-kernel Philox) : %.1f us" % timed(lambda: h.rollout(s0, native.SAMPLE_GAUSSIAN, 1, 0, d_mu=mu, d_sd=sd)))
print("uniform  (in-kernel Philox) : %.1f us" % timed(lambda: h.rollout(s0, native.SAMPLE_UNIFORM, 1, 0)))
print("injected actions (loads)    : %.1f us" % timed(lambda: h.rollout(s0, native.SAMPLE_INJECT_ACTIONS, d_injected=inj)))
