"""Fixed costs of the tcgen05 rollout kernel outside its step loop (cfg 3): launch-to-launch time per rollout
vs the kernel's own clock64() stamps of the first and last steps."""
import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from mbrl_b200 import native
from mbrl_b200.synthetic import synthetic_problem, synthetic_state
O, A, U, H, N = 17, 6, 200, 30, 16384
prob = synthetic_problem(O, A, U)
h = native.NativePlanner(O, A, U, H, N, engine="fp16")
h.load_problem(prob)
s0 = synthetic_state(prob, 0)[None].cuda()
mu = torch.zeros(1, H, A, device="cuda"); sd = torch.ones(1, H, A, device="cuda")
for _ in range(3):
    h.rollout(s0, native.SAMPLE_GAUSSIAN, 1, 0, d_mu=mu, d_sd=sd)
torch.cuda.synchronize()
ts=[]
for i in range(5):
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); h.rollout(s0, native.SAMPLE_GAUSSIAN, 1, i, d_mu=mu, d_sd=sd); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e3)
print("kernel us (events, single launch incl. launch overhead):", sorted(ts))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(20): h.rollout(s0, native.SAMPLE_GAUSSIAN, 1, i, d_mu=mu, d_sd=sd)
g.replay(); torch.cuda.synchronize()
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
a.record(); g.replay(); b.record(); torch.cuda.synchronize(); print("graph of 20 rollouts: us per rollout", a.elapsed_time(b)*1e3/20)
h.tc_debug(True)
h.rollout(s0, native.SAMPLE_GAUSSIAN, 1, 0, d_mu=mu, d_sd=sd)
torch.cuda.synchronize()
h.tc_debug(True, fetch=True)
t = h.tc_timeline[:H+2].astype(np.float64)
t0 = t[t > 0].min()
np.set_printoptions(linewidth=250, suppress=True)
for r in (0,1,2,H-2,H-1,H,H+1):
    print(r, np.where(t[r,:26] > 0, t[r,:26]-t0, -1).astype(np.int64))
