"""Per-step critical-path timeline of the tcgen05 rollout kernel (tile 1), from the kernel's
own clock64() stamps.  Run on the GPU box:  python profiles/tc_timeline.py [engine]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mbrl_b200 import native
from mbrl_b200.synthetic import synthetic_problem, synthetic_state

eng = sys.argv[1] if len(sys.argv) > 1 else "fp16"
O, A, U, H, N = 17, 6, 200, 30, 16384
prob = synthetic_problem(O, A, U)
h = native.NativePlanner(O, A, U, H, N, engine=eng)
h.load_problem(prob)
s0 = synthetic_state(prob, 0)[None].cuda()
mu = torch.zeros(1, H, A, device="cuda"); sd = torch.ones(1, H, A, device="cuda")
for _ in range(3):
    h.rollout(s0, native.SAMPLE_GAUSSIAN, 1, 0, d_mu=mu, d_sd=sd)
h.tc_debug(True)
h.rollout(s0, native.SAMPLE_GAUSSIAN, 1, 0, d_mu=mu, d_sd=sd)
torch.cuda.synchronize()
h.tc_debug(True, fetch=True)
t = h.tc_timeline[:H].astype(np.float64)
names = {8: "mma:GEMM-B all issued", 9: "mma:dB commit issued", 10: "mma:y(h-1) consumed seen", 0: "mma:step start (dA committed)", 1: "mma:GEMM-B chunk0 released", 2: "mma:GEMM-B last chunk released", 3: "mma:GEMM-A y+xa ready",
         16: "mma:GEMM-A chunk0 released", 17: "mma:GEMM-A last chunk released", 4: "epi:dA ready", 5: "epi:epiA done",
         6: "epi:dB ready", 7: "epi:epiB done", 18: "mma:before wait GEMM-A chunk0", 19: "epi:warp0 first unit A", 20: "epi:warp0 first unit B", 21: "epi:warp15 first unit A", 22: "epi:warp15 first unit B", 23: "epi:slowest warp first unit A", 24: "epi:slowest warp first unit B", 12: "smp:dA ready", 13: "smp:actions(h+1) arrived", 14: "cost:dA ready", 15: "cost:y consumed"}
base = t[:, 0:1]
rel = t - base
print("median cycles since mma step start, steps 2..H-2 (step period = %d)" % np.median(np.diff(t[2:-1, 0])))
for e in sorted(names, key=lambda e: np.median(rel[2:-1, e])):
    print(f"  {names[e]:30s} {np.median(rel[2:-1, e]):8.0f}")
