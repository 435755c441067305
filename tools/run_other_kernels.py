#!/usr/bin/env python3
"""A short run of the kernels that are not on the default bench loop (for one ncu capture each): the generic step kernel on
the XL stress configuration, the r1 macro-tile kernel (`tuning fast_impl=trip`), the device maze reset follow-up launch,
the ring-cache rebuild and reset_all."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_env_b200 import PRESETS, PlantOSVecEnv

def steps(env, n, k=12):
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    for _ in range(k):
        env.step_async(torch.randint(0, 5, (n,), device="cuda", generator=g)); env.step_wait()
    env.check()

n = 32768
env = PlantOSVecEnv(n, device="cuda:0", seed=0, full_infos=False, max_steps=100, **PRESETS["xl"]); env.reset()
gid = torch.arange(n, device="cuda", dtype=torch.int64)
env.set_state(scalars={"step_count": (((gid * 2654435761) % 4294967296) % 100).to(torch.int32)})
steps(env, n); print("xl:", env.last_step_kernel); env.close()
n = 131072
env = PlantOSVecEnv(n, device="cuda:0", seed=0, full_infos=False, map_source="maze", **PRESETS["training"]); env.reset()
gid = torch.arange(n, device="cuda", dtype=torch.int64)
env.set_state(scalars={"step_count": (((gid * 2654435761) % 4294967296) % 1000).to(torch.int32)})
steps(env, n); print("maze:", env.last_step_kernel); env.close()
env = PlantOSVecEnv(n, device="cuda:0", seed=0, full_infos=False, tuning={"fast_impl": "trip"}, **PRESETS["training"]); env.reset()
steps(env, n); print("trip:", env.last_step_kernel); env.reset(); env.close()
