#!/usr/bin/env python3
"""Where the time of a SHORT timed region goes: host time of one step_many call vs the event time around it."""
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
for K in (4, 8, 16, 20, 32):
    a = acts[:K]
    for _ in range(3):
        env.step_many(a, with_flags=True)
    torch.cuda.synchronize()
    res = []
    for rep in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        t0 = time.perf_counter()
        env.step_many(a, with_flags=True)
        t1 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        res.append((e0.elapsed_time(e1) * 1e3, (t1 - t0) * 1e6))
    res.sort()
    ev, host = res[len(res) // 2]
    # back-to-back pairs: the second call is queued while the first runs
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        env.step_many(a, with_flags=True)
    e1.record(); torch.cuda.synchronize()
    print("K=%d single call: events %.1f us (%.2f us/step), host side of the call %.1f us; 8 calls back to back %.2f us/step"
          % (K, ev, ev / K, host, e0.elapsed_time(e1) * 1e3 / (8 * K)), flush=True)
env.close()
