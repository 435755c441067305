"""Throughput of env.step_many (state-resident multi-step kernel) vs K; CUDA events."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from rl_env_b200.vec_env import PlantOSVecEnv, PRESETS
N = int(os.environ.get("N", "131072"))
env = PlantOSVecEnv(N, device="cuda:0", seed=1, full_infos=False, **PRESETS["training"])
env.reset()
for K in (4, 16, 32):
    acts = torch.randint(0, 5, (K, N), device="cuda")
    for _ in range(3): env.step_many(acts)
    torch.cuda.synchronize()
    reps = max(3, 320 // K)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): env.step_many(acts)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / (reps * K)
    print(f"step_many N={N} K={K}: {us:.2f} us/step  {N / us * 1e6:.3e} env-steps/s  frac {N * 441 / us / 1e3 / 6543.4:.3f}", flush=True)
env.check()
