#!/usr/bin/env python3
"""Step time of the other BASELINE.json configurations (secondary numbers, not the bench.py line):
CUDA-event timing of `steps` back-to-back PlantOSVecEnv steps with a ring of pre-generated actions.
usage: tools/bench_presets.py [steps]"""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_env_b200 import PlantOSVecEnv, PRESETS

def run(label, n, steps, **kw):
    env = PlantOSVecEnv(n, device="cuda:0", seed=0, obs_ring=5, full_infos=False, **kw)
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    acts = [torch.randint(0, 5, (n,), device="cuda", generator=g) for _ in range(16)]
    for i in range(50): env.step(acts[i % 16])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps): env.step_async(acts[i % 16])
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / steps
    d = env.obs_dim
    out = {"config": label, "envs": n, "kernel": env.kernel_name, "obs_dim": d, "us_per_step": round(us, 2),
           "env_steps_per_s": round(n / us * 1e6), "alg_GBps": round(n * (4 * d + 13) / us / 1e3, 1)}
    print(json.dumps(out)); env.close()

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
run("training preset (bench.py workload)", 131072, steps, **PRESETS["training"])
run("ctor-default preset G21/P8/O50/R2/C10", 131072, steps, **PRESETS["default"])
run("configs[2]: 4096 envs, training preset", 4096, steps, **PRESETS["training"])
run("configs[2]: 4096 envs, ctor-default preset", 4096, steps, **PRESETS["default"])
run("configs[4]: XL stress G64/P64/O600/R32/C16, max_steps 100", 32768, max(100, steps // 3), max_steps=100, **PRESETS["xl"])
run("training preset + CurriculumWrapper 'a2c'", 131072, max(100, steps // 3), curriculum="a2c", **PRESETS["training"])

# the same small-batch configurations as one CUDA-graph launch of 50 steps (make_rollout)
def run_graph(label, n, k, reps, **kw):
    env = PlantOSVecEnv(n, device="cuda:0", seed=0, full_infos=False, **kw)
    env.reset()
    roll = env.make_rollout(k)
    acts = torch.randint(0, 5, (k, n), device="cuda")
    for _ in range(3): roll(acts)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): roll.graph.replay()
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / (reps * k)
    print(json.dumps({"config": label + f", CUDA graph of {k} steps", "envs": n, "kernel": env.kernel_name,
                      "us_per_step": round(us, 2), "env_steps_per_s": round(n / us * 1e6)})); env.close()

run_graph("configs[2]: 4096 envs, training preset", 4096, 50, 20, **PRESETS["training"])
run_graph("configs[2]: 4096 envs, ctor-default preset", 4096, 50, 20, **PRESETS["default"])
run_graph("training preset", 131072, 50, 10, **PRESETS["training"])
