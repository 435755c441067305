"""Per-step time of pipelined graph rollouts as a function of the graph length K (per-replay overhead)."""
import os, sys, torch, time
sys.path.insert(0, os.getcwd())
from rl_env_b200.vec_env import PlantOSVecEnv, PRESETS
N = 131072
for pipe in (True, False):
    for K in (16, 48, 128):
        env = PlantOSVecEnv(N, device="cuda:0", seed=1, obs_ring=2, full_infos=False, **PRESETS["training"])
        env.reset()
        roll = env.make_rollout(K, pipelined=pipe)
        roll.actions.copy_(torch.randint(0, 5, (K, N), device="cuda"))
        for _ in range(3): roll.graph.replay()
        torch.cuda.synchronize()
        reps = max(4, 640 // K)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for _ in range(reps): roll.graph.replay()
        b.record(); torch.cuda.synchronize()
        host = (time.perf_counter() - t0) * 1e6 / (reps * K)
        print(f"pipelined={pipe} K={K} reps={reps}: {a.elapsed_time(b) * 1e3 / (reps * K):.2f} us/step (host wall {host:.2f})", flush=True)
        env.check(); env.close(); del roll, env
        torch.cuda.empty_cache()
