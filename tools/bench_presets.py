#!/usr/bin/env python3
"""Step time of the other BASELINE.json configurations and presets (secondary numbers, not the bench.py line).
For every configuration: (a) eager single-step launches (closed-loop style: full dependency between steps),
(b) env.step_many with 16 steps per call (one launch of the state-resident kernel where it exists; consecutive
launches pipelined on the device, plantos_set_pipelining).
CUDA-event timing, staggered episode phases (hash(env id) mod max_steps) unless a curriculum is active.
usage: tools/bench_presets.py [steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_env_b200 import PRESETS, PlantOSVecEnv  # noqa: E402


def run(label, n, steps, **kw):
    env = PlantOSVecEnv(n, device="cuda:0", seed=0, obs_ring=5, full_infos=False, **kw)
    env.reset()
    if kw.get("curriculum") is None:
        gid = torch.arange(n, device="cuda", dtype=torch.int64)
        env.set_state(scalars={"step_count": (((gid * 2654435761) % 4294967296) % env.max_steps).to(torch.int32)})
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    acts = torch.randint(0, 5, (16, n), device="cuda", generator=g)
    d = env.obs_dim
    out = {"config": label, "envs": n, "obs_dim": d}
    for mode in ("step", "step_many16"):
        env.set_pipelining(mode == "step_many16")      # consecutive rollout launches overlap on the device; single steps: plain
        def once():
            if mode == "step":
                for t in range(16):
                    env.step_async(acts[t])
            else:
                env.step_many(acts)
        for _ in range(3):
            once()
        torch.cuda.synchronize()
        reps = max(2, steps // 16)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            once()
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / (reps * 16)
        out[mode] = {"kernel": env.last_step_kernel, "us_per_step": round(us, 2), "env_steps_per_s": round(n / us * 1e6),
                     "alg_GBps": round(n * (4 * d + 13) / us / 1e3, 1), "frac_of_6543": round(n * (4 * d + 13) / us / 1e3 / 6543.4, 3)}
    env.check()
    print(json.dumps(out), flush=True)
    env.close()
    del env
    torch.cuda.empty_cache()


steps = int(sys.argv[1]) if len(sys.argv) > 1 else 480
T, DF, XL = PRESETS["training"], PRESETS["default"], PRESETS["xl"]
run("training preset, 131072 envs (bench.py workload)", 131072, steps, **T)
run("ctor-default preset G21/P8/O50/R2/C10, 131072 envs", 131072, steps, **DF)
run("configs[2]: 4096 envs, training preset", 4096, steps, **T)
run("configs[2]: 4096 envs, ctor-default preset", 4096, steps, **DF)
run("configs[4]: XL stress G64/P64/O600/R32/C16, 32768 envs, max_steps 100", 32768, max(96, steps // 3), max_steps=100, **XL)
run("training preset + CurriculumWrapper 'a2c', 131072 envs", 131072, max(96, steps // 3), curriculum="a2c", **T)
run("training preset, map_source='maze' (device maze generator), 131072 envs", 131072, max(96, steps // 3), map_source="maze", **T)
run("training preset, 1048576 envs on one GPU", 1048576, max(96, steps // 3), **T)
