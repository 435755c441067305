#!/usr/bin/env python3
"""bench.py -- env-steps/s of the PlantOS env-step hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one VecEnv.step over all envs (transition + LIDAR observation + auto-reset, fused).
Workload (BASELINE.json configs[3], the one the metric is quoted on): training preset
G25/P10/O12/R6/C16 (D=107), 131 072 envs per GPU -- 1 048 576 envs on 8 GPUs -- weak scaling,
Philox maps, i.i.d. uniform actions; the envs start at staggered episode phases
(step_count = hash(env id) mod max_steps), so about N / 1000 envs truncate and are regenerated in EVERY
step of the timed region.

Timed loop (`--loop`): `rollout` (default) = env.step_many, `--chunk` (32) steps per plantos_rollout call with
pre-generated actions: every step reads its own action vector and writes its own observation / reward /
done buffers, the K steps of a call are one launch of the state-resident kernel, consecutive launches
pipelined on the device (a --steps that is not a multiple of the chunk rides on the last call); `graph` = a replayed
CUDA graph of 16 single-step plantos_step launches (round 1's loop; consecutive launches pipelined on
the device unless --no-pipeline); `eager` = one plantos_step call per step.  The line always carries
the single-launch-per-step numbers next to the headline (`step_launch`).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA events, max over
ranks); `e2e` = the same metric through plantos_step_host with pinned HOST buffers (H2D
actions, D2H obs/reward/done inside the timed region); `roofline` = algorithmic bytes
(441 B per env-step at D=107) over the measured step time against the measured
HBM copy peak; `cpu_baseline` = the oracle port of the reference on this box's host cores.
`--impl reference` times that CPU implementation alone.
"""
from __future__ import annotations

import argparse
import gc
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ENVS_PER_GPU = 131072
# --preset: "training" is the headline workload (BASELINE.json configs[3]); "default" = the reference ctor's
# defaults (configs[0]'s env at batch size); "xl" = configs[4], the 64x64 / range-32 stress map with a 100-step
# horizon (reset-heavy), 32 768 envs per GPU
PRESETS = {
    "training": (dict(grid_size=25, num_plants=10, num_obstacles=12, lidar_range=6, lidar_channels=16), 131072, 1000,
                 "PlantOS training preset G25/P10/O12/R6/C16"),
    "default": (dict(grid_size=21, num_plants=8, num_obstacles=50, lidar_range=2, lidar_channels=10), 131072, 1000,
                "PlantOS ctor-default preset G21/P8/O50/R2/C10"),
    "xl": (dict(grid_size=64, num_plants=64, num_obstacles=600, lidar_range=32, lidar_channels=16), 32768, 100,
           "PlantOS XL stress preset G64/P64/O600/R32/C16, max_steps 100"),
}
PRESET, _, MAX_STEPS, PRESET_LABEL = PRESETS["training"]
OBS_DIM = 5 * PRESET["lidar_channels"] + 27
B_ALG = 4 * OBS_DIM + 4 + 1 + 8          # obs f32[D] + reward f32 + done u8 + action i64


def select_preset(name: str, envs_per_gpu):
    global PRESET, ENVS_PER_GPU, MAX_STEPS, PRESET_LABEL, OBS_DIM, B_ALG
    PRESET, default_envs, MAX_STEPS, PRESET_LABEL = PRESETS[name]
    ENVS_PER_GPU = int(envs_per_gpu) if envs_per_gpu else default_envs
    OBS_DIM = 5 * PRESET["lidar_channels"] + 27
    B_ALG = 4 * OBS_DIM + 4 + 1 + 8
PREWARM_CHUNKS = 40                       # untimed steps (x ACTION_RING) before the W warm-up steps: the GPU has idled for seconds
                                          # while the process started, and 200 ms of idle already cost a 20-step region 40 %
ACTION_RING = 16                          # steps per rollout call / per replayed graph
ACTION_POOL = 4                           # distinct [16, N] action blocks the timed loop cycles through (64 i.i.d. vectors)
OBS_RING = 5                              # rollout-buffer depth (A2C n_steps=5, A2C_training.py:229-247)
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"


def workload_name(n_gpus: int) -> str:
    return (f"{PRESET_LABEL} (D={OBS_DIM}), {ENVS_PER_GPU} envs/GPU x {n_gpus} GPU "
            f"= {ENVS_PER_GPU * n_gpus} envs, fused step+LIDAR obs+auto-reset, random actions, staggered episode phases")


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(envs_per_gpu: int, kernel: str, preset: str):
    """Steady-state DRAM bytes per STEP of the dominant kernel from the committed ncu capture
    (profiles/r2_traffic.json: application replay, no cache flush); only valid for the workload, the
    kernel and the kernel CODE it was captured on (rl_env_b200.build.kernel_source_hash: comments do not count), else null."""
    try:
        from rl_env_b200.build import kernel_source_hash
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            t = json.load(f)
        if (envs_per_gpu == int(t.get("envs", 0)) and preset == t.get("preset") and kernel in t.get("kernels", {})
                and kernel_source_hash() == t.get("kernel_source_sha")):
            return int(t["kernels"][kernel]["dram_bytes_per_step"]), t["kernels"][kernel].get("source", "")
    except Exception:
        pass
    return None, ""


# ------------------------------------------------------------------ CPU baseline (oracle port)
def _cpu_worker(conn, n_envs, seed, literal_trig):
    import random
    import numpy as np
    from oracle.plantos_oracle import OracleVecEnv
    random.seed(seed)
    rng = np.random.default_rng(seed)
    env = OracleVecEnv(n_envs, literal_trig=literal_trig, max_steps=MAX_STEPS, **PRESET)
    env.reset()
    conn.send("ready")
    while True:
        msg = conn.recv()
        if msg is None:
            break
        for _ in range(msg):  # msg = number of VecEnv steps to run
            env.step(rng.integers(0, 5, size=n_envs))
        conn.send(msg * n_envs)
    conn.close()


class CpuPool:
    """One process per host core, each stepping its own slice of envs through the oracle's
    DummyVecEnv+Monitor restatement (SubprocVecEnv's work without its pipe traffic: an
    upper bound on what SB3's SubprocVecEnv could reach on these cores)."""

    def __init__(self, envs_per_worker: int, workers: int | None = None):
        self.workers = workers or (os.cpu_count() or 1)
        self.envs_per_worker = envs_per_worker
        ctx = mp.get_context("fork")
        self.conns, self.procs = [], []
        for w in range(self.workers):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_cpu_worker, args=(b, envs_per_worker, 1000 + w, True), daemon=True)
            p.start()
            self.conns.append(a)
            self.procs.append(p)
        for c in self.conns:
            assert c.recv() == "ready"

    @property
    def num_envs(self) -> int:
        return self.workers * self.envs_per_worker

    def run(self, vec_steps: int) -> int:
        for c in self.conns:
            c.send(vec_steps)
        return sum(c.recv() for c in self.conns)

    def close(self):
        for c in self.conns:
            try:
                c.send(None)
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=5)


def cpu_baseline(seconds: float = 12.0):
    pool = CpuPool(envs_per_worker=REF_ENVS_PER_WORKER)
    pool.run(20)  # warm-up
    t0 = time.perf_counter()
    steps = 0
    chunk = 10
    while time.perf_counter() - t0 < seconds:
        steps += pool.run(chunk)
    dt = time.perf_counter() - t0
    cores = pool.workers
    n_envs = pool.num_envs
    pool.close()
    return {"value": steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} env-steps in {dt:.1f} s: {n_envs} envs ({REF_ENVS_PER_WORKER} per process, {cores} processes, "
                      f"no IPC on the step path) of the same preset, oracle/plantos_oracle.py"}


REF_ENVS_PER_WORKER = 64      # fixed sample: 64 envs per host core
REF_SECONDS = float(os.environ.get("PLANTOS_BENCH_REF_SECONDS", "10.0"))            # CPU time the K timed steps of the reference arm add up to


def run_reference(args, rank: int):
    """`--impl reference`: the reference's CPU implementation (oracle port) on all host cores.  The
    sample is FIXED (64 envs per core); one bench "step" = `reps` VecEnv steps over it, with reps
    calibrated so that the K timed steps take about REF_SECONDS (short timings were +-40 % noisy)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    pool = CpuPool(envs_per_worker=REF_ENVS_PER_WORKER, workers=cores)
    n_envs = pool.num_envs
    pool.run(3)                                   # warm the interpreters
    t0 = time.perf_counter()
    done = pool.run(10)
    rate = done / (time.perf_counter() - t0)      # env-steps/s, calibration only
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    reps = max(1, int(round(REF_SECONDS * rate / (steps * n_envs))))
    pool.run(min(warmup * reps, max(1, int(2.0 * rate / n_envs))))      # W warm-up steps, at most ~2 s
    t0 = time.perf_counter()
    n = 0
    for _ in range(steps):
        n += pool.run(reps)
    dt = time.perf_counter() - t0
    pool.close()
    value = n / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "python float64 / int", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus), "sample_envs": n_envs, "vec_steps_per_bench_step": reps,
                   "note": "CPU port of plantos_env.py (oracle/plantos_oracle.py), one process per host core, no IPC "
                           "on the step path; each bench step is `vec_steps_per_bench_step` VecEnv steps over the "
                           "fixed sample (64 envs per core)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} env-steps in {dt:.1f} s over {n_envs} envs ({REF_ENVS_PER_WORKER} per core)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons read through NVML while the timed steps execute.

    The timed region is enqueued first (kernel launches are asynchronous), then the calling thread polls NVML
    until the closing event has completed: every sample is taken while the GPU is running timed steps, and no
    NVML call competes with the launches for the driver (a sampler thread that started polling next to the
    first launch cost 50-70 us of a 300 us region)."""

    def __init__(self, index: int):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._h = None

    def sample(self):
        if self._h is None:
            return
        nv = self._nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
            try:
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
            except Exception:
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                     0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}
            for bit, name in names.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def sample_until(self, event, max_samples: int = 4000):
        """Poll until `event` (recorded after the last timed step) has completed; at least one sample."""
        self.sample()
        while not event.query() and len(self.samples) < max_samples:
            self.sample()
            time.sleep(0.0005)
        self.in_flight = len(self.samples)
        self.sample()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "samples_while_timed_steps_ran": getattr(self, "in_flight", 0)}


# ------------------------------------------------------------------ ours
def run_ours(args, rank: int, world: int, local_rank: int):
    # stdout carries exactly one JSON line: anything libraries write to fd 1 meanwhile (NCCL's version
    # banner, for one) goes to stderr instead
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from rl_env_b200 import make_sharded

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # the rank's thread and its pinned host buffers go to the NUMA node of its GPU (matters for the
    # host-buffer e2e path when several ranks copy 57 MB per step at the same time)
    from rl_env_b200.affinity import bind_to_gpu
    placement = bind_to_gpu(local_rank) if not args.no_numa_bind else {"numa_node": None, "cpus": None, "mempolicy": False}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.envs_per_gpu
    env = make_sharded(n * world, rank, world, local_device=local_rank, seed=0, kernel=args.kernel,
                       obs_ring=OBS_RING, track_terminal_obs=not args.no_terminal_obs, full_infos=False,
                       max_steps=MAX_STEPS, **PRESET)
    assert env.num_envs == n and env.obs_dim == OBS_DIM
    gen = torch.Generator(device=dev).manual_seed(rank)
    # ACTION_POOL blocks of ACTION_RING i.i.d. action vectors (the timed loop cycles through the blocks)
    pool = [torch.randint(0, 5, (ACTION_RING, n), device=dev, dtype=torch.int64, generator=gen) for _ in range(ACTION_POOL)]
    env.reset()
    # staggered episode phases: env i starts with step_count = global id mod max_steps, so ~N / max_steps
    # envs hit the 1000-step truncation (terminal observation, Philox map, fresh observation) in EVERY step
    if not args.no_stagger:
        # (a multiplicative hash of the global env id: neighbouring envs must not truncate in neighbouring
        # steps, or the 32 envs of one tile would reset 32 steps in a row)
        gid = torch.arange(n, device=dev, dtype=torch.int64) + env.env_id_base
        phase = ((gid * 2654435761) % 4294967296) % env.max_steps
        env.set_state(scalars={"step_count": phase.to(torch.int32)})
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pipelined = not args.no_pipeline
    loop = args.loop
    roll = None
    env.set_pipelining(pipelined)          # every loop: eager steps (the obs ring gives every step its own buffer), graphs, rollout calls
    if loop == "graph":
        roll = env.make_rollout(ACTION_RING, with_flags=True, pipelined=pipelined)

    # the optional NCCL all-reduce of the 8 episode statistics lands INSIDE the timed window whatever K is
    stats_every = 0
    if world > 1 and args.stats_every:
        stats_every = max(ACTION_RING, min(args.stats_every, max(1, args.steps // 2)) // ACTION_RING * ACTION_RING)
    counters = {"stats": 0, "launches": 0}

    # chunks: `--chunk` steps per rollout call (ACTION_RING per graph replay); a remainder shorter than the chunk
    # rides on the last call (rollout loop: one call of up to 2 * chunk - 1 steps), so that a short timed region
    # (--steps 20) is ONE call, not a 16-step call plus a 4-step call with a host round trip in between
    pool2 = [torch.cat([pool[(b + j) % ACTION_POOL] for j in range(4)]) for b in range(ACTION_POOL)]   # [4 * ACTION_RING, n] each

    chunk = min(max(args.chunk, 1), 2 * ACTION_RING) if loop == "rollout" else ACTION_RING

    pending_stats = []

    def plan_steps(k, start):
        """The calls of k consecutive steps as (m, action block) pairs; `start` only selects which pre-generated
        action blocks are used.  Built before the clock starts: slicing tensors is not part of a step."""
        plan, i = [], 0
        while i < k:
            m = k - i if k - i < 2 * chunk else chunk
            if loop == "graph":
                m = min(m, ACTION_RING) if m >= ACTION_RING else m
            b = ((start + i) // ACTION_RING) % ACTION_POOL
            if loop == "rollout":
                plan.append((m, pool2[b][:m]))
            elif loop == "graph" and m >= ACTION_RING:
                plan.append((ACTION_RING, pool[b]))
            else:
                plan.append((m, [pool2[b][t] for t in range(m)]))
            i += plan[-1][0]
        return plan

    def run_steps(k, start, plan=None):
        for m, acts in (plan if plan is not None else plan_steps(k, start)):
            done_steps = counters.get("steps", 0)
            if loop == "rollout":
                env.step_many(acts, with_flags=True)                 # ONE plantos_rollout call: m steps
                counters["launches"] += 1
            elif loop == "graph" and m == ACTION_RING and not isinstance(acts, list):
                roll.actions.copy_(acts, non_blocking=True)          # (1 MB device copy per 16 steps, inside the timed region)
                roll.graph.replay()
                counters["launches"] += ACTION_RING
            else:
                for a in acts:
                    env.step_async(a)
                    env.step_wait()
                counters["launches"] += m
            counters["steps"] = done_steps + m
            # episode statistics: whenever the steps just enqueued crossed a multiple of `stats_every`, the counters
            # are snapshotted behind them and all-reduced (NCCL, 8 doubles) on the process group's stream, next to the
            # steps that follow; the timed region ends only after every such all-reduce has completed.  (Issued BEFORE a
            # 20-step launch instead, the NCCL kernel did not get an SM until that launch drained: 20.6 vs 17.5 us/step
            # at 8 GPUs.)
            if stats_every and (done_steps + m) // stats_every > done_steps // stats_every:
                pending_stats.append(env.episode_stats_tensor(all_reduce=True, async_op=True))
                counters["stats"] += 1
        for ps in pending_stats:                                     # (a stream-level wait, no host sync)
            ps.wait()
        pending_stats.clear()

    # untimed extra steps before the W warm-up steps: first launches, graph instantiation and the
    # first NCCL collective on every rank (a cold rank once made a whole 4-GPU run 30 % slower)
    sampler = ClockSampler(local_rank)      # (NVML initialised here: tens of ms of idle GPU right before the timed
    sampler.sample()                        #  region drop its clocks -- a 300 us region then runs 40 % slower)
    sampler.samples.clear()
    # everything the host can do ahead of the clock happens BEFORE the untimed steps: anything slow between them
    # and the opening event would let the GPU idle (a GPU that idled for tens of ms runs the first launch slower)
    timed_plan = plan_steps(args.steps, 2 * ACTION_RING + args.warmup)
    gc.collect()
    gc.disable()                                           # (no collector pause between the opening event and the launches)
    run_steps(PREWARM_CHUNKS * ACTION_RING, 0)
    if args.steps % chunk:                                # (the odd-sized last chunk of the timed region: its buffers exist now)
        run_steps(args.steps % chunk + (chunk if args.steps > chunk else 0), 0)
    if world > 1:
        env.episode_stats_tensor(all_reduce=True)
    barrier()
    run_steps(args.warmup, 2 * ACTION_RING)
    barrier()
    counters["stats"] = counters["launches"] = counters["steps"] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_steps(args.steps, 0, timed_plan)
    e1.record()
    gc.enable()
    sampler.sample_until(e1)
    barrier()
    ms = e0.elapsed_time(e1)
    launches, stats_in_window = counters["launches"], counters["stats"]
    kernel_of_loop = env.last_step_kernel
    env.check()

    def timed(fn, reps):
        fn(); fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e3 / reps

    # the single-launch-per-step numbers, always reported next to the headline: (a) one launch alone between
    # two events (isolated duration), (b) 16-step graphs of plantos_step launches, pipelined and with a
    # grid-wide dependency between consecutive launches
    step_launch = {}
    if not args.no_step_launch:
        env.set_pipelining(False)
        iso = []
        for i in range(64):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            env.step_async(pool[0][i % ACTION_RING])
            b.record()
            env.step_wait()
            iso.append((a, b))
        torch.cuda.synchronize()
        iso_us = sorted(a.elapsed_time(b) * 1e3 for a, b in iso)
        step_launch["isolated_launch_us_median"] = iso_us[len(iso_us) // 2]
        for label, pl in (("graph16_pipelined_us_per_step", True), ("graph16_plain_us_per_step", False)):
            g = env.make_rollout(ACTION_RING, with_flags=True, pipelined=pl)
            g.actions.copy_(pool[1])
            step_launch[label] = timed(g.graph.replay, 12) / ACTION_RING
            del g
        step_launch["kernel"] = env.last_step_kernel
        step_launch["frac_pipelined"] = n * B_ALG / step_launch["graph16_pipelined_us_per_step"] / 1e3 / measured_peak_gbs()[0]
        step_launch["frac_plain"] = n * B_ALG / step_launch["graph16_plain_us_per_step"] / 1e3 / measured_peak_gbs()[0]

    # end to end through the host-buffer entry point (plantos_step_host): numpy in / numpy out, pinned host
    # buffers, H2D actions + D2H obs / reward / done inside the timed region, one call per step
    host_actions = [pool[0][i].cpu().numpy() for i in range(4)]
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    for i in range(3):
        env.step_host(host_actions[i % 4])
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        env.step_host(host_actions[i % 4])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    stats = env.episode_stats_tensor(all_reduce=True).cpu().tolist()
    t_ms = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    per_rank_ms, per_rank_e2e = [ms], [e2e_s * 1e3]
    if world > 1:
        gathered = [torch.zeros_like(t_ms) for _ in range(world)]
        dist.all_gather(gathered, t_ms)
        per_rank_ms = [float(g[0]) for g in gathered]     # the timed region of every rank (value uses the max)
        per_rank_e2e = [float(g[1]) for g in gathered]
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms, e2e_ms = t_ms.tolist()
    kernel_name = env.kernel_name
    state_bytes = env.state_bytes_per_env
    env.close()

    if rank == 0:
        total_envs = n * world
        value = total_envs * args.steps / (ms * 1e-3)
        peak, peak_src = measured_peak_gbs()
        step_s = ms * 1e-3 / args.steps
        achieved = n * B_ALG / step_s / 1e9
        traffic, traffic_src = measured_traffic(n, kernel_of_loop, args.preset)
        launch_desc = {
            "rollout": "env.step_many: %d steps in %d plantos_rollout call(s) (%d steps per call, a remainder rides on the last; one launch of the "
                       "state-resident kernel %s per call; every step reads its own action vector and writes its own obs/reward/done buffers)"
                       % (args.steps, launches, chunk, kernel_of_loop)
                       + ("; consecutive launches pipelined on the device (plantos_set_pipelining)" if pipelined else ""),
            "graph": "CUDA graph of %d single-step plantos_step launches, replayed" % ACTION_RING
                     + ("; consecutive launches pipelined on the device (plantos_set_pipelining)" if pipelined else ""),
            "eager": "eager, one plantos_step launch per step" + ("; pipelined (plantos_set_pipelining)" if pipelined else ""),
        }[loop]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "int (2-bit cells, 4-bit/u16 visit counts) + f32 table-driven obs/reward", "data": "synthetic",
            "config": {"workload": workload_name(world), "envs_per_gpu": n, "obs_dim": OBS_DIM,
                       "kernel": kernel_name, "loop_kernel": kernel_of_loop, "maps": "philox seed 0", "state_bytes_per_env": state_bytes,
                       "actions": f"{ACTION_POOL * ACTION_RING} pre-generated i.i.d. uniform vectors, cycled",
                       "episode_phases": "all envs start at step 0" if args.no_stagger else f"staggered: step_count = hash(env id) mod {MAX_STEPS} (about N/{MAX_STEPS} auto-resets in every step)",
                       "l2": f"outputs larger than L2: {(chunk if loop == 'rollout' else ACTION_RING) if loop != 'eager' else OBS_RING} observation buffers x {n * OBS_DIM * 4 / 1e6:.0f} MB "
                             f"written round-robin vs 126 MB L2 (per-GPU state {n * state_bytes / 1e6:.0f} MB); no explicit flush",
                       "launch": launch_desc,
                       "untimed_steps_before_warmup": PREWARM_CHUNKS * ACTION_RING,
                       "stats_allreduce_every": stats_every, "stats_allreduces_in_timed_window": stats_in_window},
            "e2e": {"value": total_envs * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * (4 * OBS_DIM + 4 + 1),
                    "steps": e2e_steps, "api": "plantos_step_host (pinned host buffers), one call per step and GPU; bytes are per GPU, "
                                               "value is the aggregate over all GPUs (max over ranks of the wall time)",
                    "ms_per_step_by_rank": [round(m / e2e_steps, 4) for m in per_rank_e2e],
                    "host_placement_rank0": placement},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "traffic_unit": "steady-state DRAM bytes per step (ncu dram read+write at 16 steps per launch, " + traffic_src + ")" if traffic else None,
                         "peak_source": peak_src, "basis": "algorithmic bytes (obs + reward + done + action = 441 B per env-step)",
                         "alg_bytes_per_env_step": B_ALG, "kernel": kernel_of_loop,
                         "steps_per_launch": args.steps / max(1, launches),
                         "avg_launch_us": ms * 1e3 / max(1, launches)},
            "step_launch": step_launch,
            "clocks": sampler.summary(),
            "ms_per_step_by_rank": [round(m / args.steps, 6) for m in per_rank_ms],
            "episode_stats": dict(zip(("episodes", "return_sum", "length_sum", "exploration_pct_sum",
                                       "collisions_sum", "watered_sum", "terminated", "truncated"), stats)),
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.cpu_seconds)
        else:
            line["cpu_baseline"] = None
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # (NCCL_DEBUG is left as the caller set it: run_ours moves fd 1 to stderr while libraries may print)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preset", default="training", choices=sorted(PRESETS),
                    help="training (headline, configs[3]) | default (ctor defaults) | xl (configs[4]: 64x64, range 32, 32768 envs/GPU)")
    ap.add_argument("--envs-per-gpu", type=int, default=0, help="default: 131072 (32768 for --preset xl); 4096 = configs[2]")
    ap.add_argument("--kernel", default="auto", choices=["auto", "generic", "fast"])
    ap.add_argument("--stats-every", type=int, default=100)
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-terminal-obs", action="store_true")
    ap.add_argument("--loop", default="rollout", choices=["rollout", "graph", "eager"],
                    help="timed loop: step_many (--chunk steps per rollout launch), replayed graph of 16 step launches, or eager steps")
    ap.add_argument("--chunk", type=int, default=2 * ACTION_RING, help="rollout loop: steps per plantos_rollout call (<= %d)" % (2 * ACTION_RING))
    ap.add_argument("--no-graph", action="store_true", help="same as --loop eager")
    ap.add_argument("--no-pipeline", action="store_true", help="full grid-wide dependency between consecutive step launches (graph / eager loops)")
    ap.add_argument("--no-stagger", action="store_true", help="all envs start at step 0 (no auto-reset before step 1000)")
    ap.add_argument("--no-step-launch", action="store_true", help="skip the single-launch-per-step side measurements")
    ap.add_argument("--no-numa-bind", action="store_true", help="leave CPU affinity and memory policy as inherited")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    select_preset(args.preset, args.envs_per_gpu)
    args.envs_per_gpu = ENVS_PER_GPU
    if args.no_graph:
        args.loop = "eager"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        print(f"note: --gpus {args.gpus} without torchrun; launch with torch.distributed.run "
              f"--nproc-per-node {args.gpus}. Running 1 GPU.", file=sys.stderr)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
