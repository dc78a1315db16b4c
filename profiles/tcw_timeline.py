"""In-kernel clock64 timeline of the weight-streaming rollout kernel (csrc/rollout_tcw.cuh, DBG build):
tile 1 of a full cfg-5 shard (128 envs x 2048 candidates, H=50, hidden 512), so that all SMs stream
weights concurrently.  Prints per-step phase durations in SM cycles.
    python profiles/tcw_timeline.py [envs] [engine]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mbrl_b200 import native
from mbrl_b200.synthetic import synthetic_problem, synthetic_state

O, A, U, H, N = 67, 21, 512, 50, 2048
E = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ENGINE = sys.argv[2] if len(sys.argv) > 2 else "fp16"
prob = synthetic_problem(O, A, U)
h = native.NativePlanner(O, A, U, H, N, E, 1, 204, ENGINE)
h.load_problem(prob)
s0 = torch.stack([synthetic_state(prob, e) for e in range(E)]).cuda()
mu = torch.zeros(E, H, A, device="cuda"); sd = torch.ones(E, H, A, device="cuda")
h.rollout(s0, native.SAMPLE_GAUSSIAN, 1, 0, d_mu=mu, d_sd=sd); torch.cuda.synchronize()
h.tc_debug(True)
h.rollout(s0, native.SAMPLE_GAUSSIAN, 1, 0, d_mu=mu, d_sd=sd); torch.cuda.synchronize()
h.tc_debug(True, fetch=True)
t = h.tc_timeline.astype(np.int64)  # [64 steps][32 events]
names = {0: "x ready", 1: "L1c0 start", 2: "L1c0 issued", 3: "L1c1 start", 4: "L1c1 issued", 5: "L2c0 start", 6: "L2c0 issued",
         7: "L2c1 start", 8: "L2c1 issued", 9: "L3 issued+epi waited", 11: "out-epi: y ready (warp 0)", 12: "out-epi: x written (warp 0)",
         13: "sampler start", 14: "sampler done", 16: "epi L1c0 start", 17: "epi L1c0 end", 18: "epi L1c1 start", 19: "epi L1c1 end",
         26: "L3 SS half issued", 27: "L3 e0 passed", 28: "L3 TS round 0 issued", 29: "L3 e1 passed", 20: "epi L2c0 start", 21: "epi L2c0 end", 22: "epi L2c1 start", 23: "epi L2c1 end"}
print("step length (x ready -> next x ready), cycles:", np.diff(t[:H, 0])[:12], "... mean", np.diff(t[:H, 0]).mean())
for hh in (5, 20, 40):
    base = t[hh, 0]
    ev = sorted((int(t[hh, e] - base), names[e]) for e in names if t[hh, e] > 0)
    print(f"--- step {hh} (cycles after x ready)")
    for c, n in ev:
        print(f"{c:8d}  {n}")

