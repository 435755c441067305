#!/usr/bin/env python3
"""Spread of a single K=20 rollout call from an idle GPU; with / without a different-K call right before."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_env_b200 import PRESETS, PlantOSVecEnv
n = 131072
env = PlantOSVecEnv(n, device="cuda:0", seed=0, obs_ring=5, full_infos=False, **PRESETS["training"])
env.reset()
gid = torch.arange(n, device="cuda", dtype=torch.int64)
env.set_state(scalars={"step_count": (((gid * 2654435761) % 4294967296) % env.max_steps).to(torch.int32)})
acts = torch.randint(0, 5, (32, n), device="cuda")
def one(K, pre=None, stats=False):
    if pre:
        env.step_many(acts[:pre], with_flags=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    env.step_many(acts[:K], with_flags=True)
    if stats:
        env.episode_stats_tensor(all_reduce=True)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3
for K in (20,):
    for _ in range(3):
        env.step_many(acts[:K], with_flags=True); env.step_many(acts[:5], with_flags=True)
    print("K=%d alone      :" % K, " ".join("%.0f" % one(K) for _ in range(12)), flush=True)
    print("K=%d after K=5  :" % K, " ".join("%.0f" % one(K, pre=5) for _ in range(12)), flush=True)
    print("K=%d + stats    :" % K, " ".join("%.0f" % one(K, stats=True) for _ in range(12)), flush=True)
    time.sleep(0.5)
    print("K=%d after sleep:" % K, " ".join("%.0f" % (time.sleep(0.2) or one(K)) for _ in range(6)), flush=True)
env.close()
