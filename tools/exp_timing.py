import os, sys, torch, numpy as np
sys.path.insert(0, os.getcwd())
from rl_env_b200.vec_env import PlantOSVecEnv, PRESETS
N = 131072
env = PlantOSVecEnv(N, device="cuda:0", seed=1, obs_ring=5, **PRESETS["training"])
env.reset()
if os.environ.get("PIPE", "0") == "1": env.set_pipelining(True)
acts = [torch.randint(0, 5, (N,), device="cuda") for _ in range(16)]
for i in range(200):
    env.step(acts[i % 16])
torch.cuda.synchronize()
r = env.returns(terminal=True).view(torch.int64).cpu().numpy()
lo = (r & 0xffffffff).astype(np.int64); hi = ((r >> 32) & 0xffffffff).astype(np.int64)
# macro tiles start at multiples of fast_q; find rows with non-zero data at k=0
T = []
q = 28
for e0 in range(0, N - 8, 32 if os.environ.get('PLANTOS_FAST_IMPL', 'tile') == 'tile' else 4):
    if lo[e0] != 0 and hi[e0] != 0 and lo[e0+4] != 0:
        T.append([lo[e0], hi[e0], lo[e0+1], hi[e0+1], lo[e0+2], hi[e0+2], lo[e0+3], hi[e0+3], lo[e0+4], hi[e0+4]])
T = np.array(T, dtype=np.int64)
os.makedirs("gpurun_out/r2", exist_ok=True)
np.save("gpurun_out/r2/stamps_%s.npy" % ("pipe" if os.environ.get("PIPE", "0") == "1" else "plain"), T)
print("warps", len(T))
base = T[:, 1].max()  # last griddep_wait return ~ previous kernel end
t0 = T[:, 0].min()
rel = T - t0
names = (["entry", "after griddep wait", "rec landed", "windows issued", "windows landed", "phase A done", "window regs built", "encode done", "expand done", "-"] if os.environ.get("PLANTOS_FAST_IMPL", "tile") == "tile" else ["entry", "after griddep wait", "tables staged", "rec landed", "target landed", "phase A done", "trip0 windows landed", "trip0 done", "all trips done", "-"])
for k in range(9):
    c = rel[:, k]
    print(f"{names[k]:24s} min {c.min():7d} p50 {int(np.median(c)):7d} p90 {int(np.percentile(c,90)):7d} max {c.max():7d} ns")
d = np.diff(T[:, :9], axis=1)
for k in range(8):
    print(f"delta {names[k]} -> {names[k+1]}: p50 {int(np.median(d[:,k]))} mean {d[:,k].mean():.0f} ns")

sm = T[:, 9]
fin = rel[:, 8]; a_done = rel[:, 5]
persm = {}
for i in range(len(T)): persm.setdefault(int(sm[i]), []).append(i)
means = np.array([fin[v].mean() for k, v in sorted(persm.items())])
spread = np.array([fin[v].max() - fin[v].min() for k, v in sorted(persm.items())])
cnt = np.array([len(v) for k, v in sorted(persm.items())])
print("SMs", len(persm), "warps/SM min/max", cnt.min(), cnt.max())
print("per-SM mean finish: min %d p50 %d max %d ns; within-SM spread p50 %d max %d" % (means.min(), np.median(means), means.max(), np.median(spread), spread.max()))
order = np.argsort(means); ks = sorted(persm.keys())
print("slowest SMs:", [(ks[i], int(means[i]), int(cnt[i])) for i in order[-8:]])
print("fastest SMs:", [(ks[i], int(means[i]), int(cnt[i])) for i in order[:8]])
bdur = fin - a_done
print("B duration per warp: p10 %d p50 %d p90 %d max %d" % tuple(np.percentile(bdur, [10, 50, 90, 100])))
sm_b = np.array([bdur[v].mean() for k, v in sorted(persm.items())])
print("per-SM mean B duration min %d p50 %d max %d; start(A done) per-SM mean min %d max %d" % (sm_b.min(), np.median(sm_b), sm_b.max(), min(a_done[v].mean() for v in persm.values()), max(a_done[v].mean() for v in persm.values())))
