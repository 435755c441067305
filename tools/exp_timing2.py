"""Phase stamps of the last FOUR consecutive k_step_tile launches (build with -DPLANTOS_EXP_TIMING):
per launch the entry / exit time distribution of its warps and the number of resident warps per SM over time.
usage: PLANTOS_LIB=build/libplantos_timing.so [PIPE=1] python tools/exp_timing2.py"""
import os, sys, torch, numpy as np
sys.path.insert(0, os.getcwd())
from rl_env_b200.vec_env import PlantOSVecEnv, PRESETS
N = 131072
env = PlantOSVecEnv(N, device="cuda:0", seed=1, obs_ring=5, **PRESETS["training"])
env.reset()
pipe = os.environ.get("PIPE", "0") == "1"
if pipe: env.set_pipelining(True)
acts = [torch.randint(0, 5, (N,), device="cuda") for _ in range(16)]
roll = env.make_rollout(16, pipelined=pipe) if os.environ.get("GRAPH", "0") == "1" else None
if roll: roll.actions.copy_(torch.stack(acts))
nrep = 12 if roll else 200
ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ea.record()
for i in range(nrep):
    if roll: roll.graph.replay()
    else: env.step_async(acts[i % 16])
eb.record()
torch.cuda.synchronize()
print("event-timed: %.2f us/step" % (ea.elapsed_time(eb) * 1e3 / (nrep * (16 if roll else 1))))
r = env.returns(terminal=True).view(torch.int64).cpu().numpy()
lo = (r & 0xffffffff).astype(np.int64); hi = ((r >> 32) & 0xffffffff).astype(np.int64)
S = np.zeros((4, N // 32, 10), dtype=np.int64)
for q in range(4):
    for k in range(5):
        S[q, :, 2 * k] = lo[5 * q + k::32]; S[q, :, 2 * k + 1] = hi[5 * q + k::32]
os.makedirs("gpurun_out/r2", exist_ok=True)
np.save("gpurun_out/r2/stamps4_%s.npy" % ("pipe" if pipe else "plain"), S)
order = np.argsort([np.median(S[q, :, 0]) for q in range(4)])
t0 = S[order[0], :, 0].min()
for q in order:
    ent, fin = S[q, :, 0] - t0, S[q, :, 8] - t0
    print("launch slot %d: entry p0/p50/p100 %6d %6d %6d   exit p0/p50/p100 %6d %6d %6d   lifetime p50 %d" % (
        q, ent.min(), np.median(ent), ent.max(), fin.min(), np.median(fin), fin.max(), np.median(fin - ent)))
