"""BASELINE config 5 (humanoid-run shape, batched environments) on one GPU's shard: 128 independent
environments x 2048 candidates, H=50, hidden 512 -- environment sharding needs no collective, so
8 GPUs run 8 of these side by side.
    python profiles/exp_cfg5_humanoid.py [iterations] [engine fp16|bf16|fp32] [envs]
fp16 / bf16 run on the weight-streaming tcgen05 kernel (csrc/rollout_tcw.cuh), fp32 on the CUDA-core engine."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mbrl_b200 import native
from mbrl_b200.synthetic import synthetic_problem, synthetic_state

O, A, U, H, N = 67, 21, 512, 50, 2048
I = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ENGINE = sys.argv[2] if len(sys.argv) > 2 else "fp16"
E = int(sys.argv[3]) if len(sys.argv) > 3 else 128
k = int(0.1 * N)
prob = synthetic_problem(O, A, U)
h = native.NativePlanner(O, A, U, H, N, E, I, k, ENGINE)
h.load_problem(prob)
s0 = torch.stack([synthetic_state(prob, e) for e in range(E)]).cuda()
out_s = torch.empty(E, H, O, device="cuda"); out_a = torch.empty(E, H, A, device="cuda")
info = torch.zeros(E, 4, dtype=torch.int32, device="cuda")
def plan(seed):
    h.plan_device(s0, out_s, out_a, info, iterations=I, elites=k, mode=native.SAMPLE_GAUSSIAN, seed=seed)
plan(0); torch.cuda.synchronize()
ts = []
for i in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); plan(1 + i); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = sorted(ts)[1]
# the rollout kernel alone
mu = torch.zeros(E, H, A, device="cuda"); sd = torch.ones(E, H, A, device="cuda")
h.rollout(s0, native.SAMPLE_GAUSSIAN, 1, 0, d_mu=mu, d_sd=sd); torch.cuda.synchronize()
ks = []
for i in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); h.rollout(s0, native.SAMPLE_GAUSSIAN, 1, i, d_mu=mu, d_sd=sd); b.record(); torch.cuda.synchronize()
    ks.append(a.elapsed_time(b))
kms = sorted(ks)[1]
flops = 2 * ((O + A) * U + U * U + U * O)
print(json.dumps(dict(workload=f"humanoid-run batched: {E} envs x {N} candidates, H={H}, I={I}, hidden={U} (BASELINE configs[4], one GPU's shard)",
                      engine=ENGINE, ms_per_plan=ms, rollout_kernel_ms=kms,
                      rollout_algorithmic_tflops=E * N * H * flops / (kms * 1e-3) / 1e12, cand_steps_per_s=E * N * H * I / (ms * 1e-3),
                      algorithmic_tflops=E * N * H * I * flops / (ms * 1e-3) / 1e12)))
