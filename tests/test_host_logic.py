"""Host-side logic that needs no GPU: sharding arithmetic, the Philox mirror, presets."""
import random

import numpy as np
import pytest

from rl_env_b200 import PRESETS, shard_range


def test_shard_range_partitions_exactly():
    for total in (1, 7, 8, 1000, 1048576):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert shard_range(1048576, 3, 8) == (393216, 131072)
    with pytest.raises(ValueError):
        shard_range(8, 8, 8)


def test_presets_are_the_reference_configs():
    assert PRESETS["default"] == dict(grid_size=21, num_plants=8, num_obstacles=50, lidar_range=2, lidar_channels=10)
    assert PRESETS["training"] == dict(grid_size=25, num_plants=10, num_obstacles=12, lidar_range=6, lidar_channels=16)


def test_philox_known_answers():
    from oracle.philox_mapgen import philox4x32_10
    # Random123 kat_vectors, philox4x32 10 rounds
    assert philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_philox_mirror_builds_valid_maps_like_the_reference_generator():
    """Construction invariants of plantos_env.py:338-372 hold for the Philox generator."""
    from oracle.philox_mapgen import generate_map
    from oracle.plantos_oracle import PlantOSOracle
    kw = PRESETS["training"]
    counts = []
    for env_id in range(300):
        cells, rover = generate_map(1, env_id, 0, kw["grid_size"], kw["num_plants"], kw["num_obstacles"])
        assert (cells >= 2).sum() == kw["num_plants"]
        assert cells[rover] == 0
        assert not (cells[0] == 1).any() and not (cells[-1] == 1).any()
        assert not (cells[:, 0] == 1).any() and not (cells[:, -1] == 1).any()
        counts.append((cells == 1).sum())
    random.seed(0)
    ref = PlantOSOracle(**kw)
    ref_counts = []
    for _ in range(300):
        ref.generate_map()
        ref_counts.append(len(ref.obstacles))
    assert abs(np.mean(counts) - np.mean(ref_counts)) < 1.5   # SURVEY: mean 25.5 for this preset
    assert min(counts) >= 4 and max(counts) <= 36


def test_maze_maps_are_connected_and_well_formed():
    """Host-side maze generator (oracle.ref_maps, the Gradio fork's 'maze' algorithm): every free cell
    is reachable from the rover (the DFS carves one connected system of rooms), plant count and codes
    are right, the rover stands on an empty cell."""
    import random
    from collections import deque
    from oracle.ref_maps import make_maps
    random.seed(4)
    cells, rover = make_maps("maze", 3, 4, grid_size=25, num_plants=10, num_obstacles=12)
    assert cells.shape == (3, 4, 25, 25) and rover.shape == (3, 4, 2)
    for i in range(3):
        for e in range(4):
            c, (rx, ry) = cells[i, e], rover[i, e]
            assert c[rx, ry] == 0 and int(((c == 2) | (c == 3)).sum()) == 10 and set(np.unique(c)) <= {0, 1, 2, 3}
            free = c != 1
            assert 100 < int(free.sum()) < 25 * 25 - 24          # carved rooms, solid outer ring
            seen = np.zeros_like(free); seen[rx, ry] = True
            queue = deque([(int(rx), int(ry))])
            while queue:
                x, y = queue.popleft()
                for dx, dy in ((-1, 0), (0, 1), (1, 0), (0, -1)):
                    nx, ny = x + dx, y + dy
                    if 0 <= nx < 25 and 0 <= ny < 25 and free[nx, ny] and not seen[nx, ny]:
                        seen[nx, ny] = True; queue.append((nx, ny))
            assert seen.sum() == free.sum()


def test_philox_maze_mirror_matches_the_forks_generator_in_distribution():
    """oracle.philox_mapgen.generate_maze_map (the host mirror of the device's map_source="maze" reset) against
    the restated generator of the Gradio fork (oracle.ref_maps.maze_map == gradio-app/plantos_env_new.py:408-604,
    pinned to the fork in test_oracle_vs_reference.py): same construction, different random stream -> same
    statistics.  Free-cell count (mean and spread), per-cell free marginal, solid outer ring, 100 % connectivity,
    plant count and thirsty fraction."""
    import random
    from collections import deque
    from oracle.philox_mapgen import generate_maze_map
    from oracle.ref_maps import maze_map
    g, n = 25, 300
    random.seed(11)
    ref = [maze_map(g, 10, 12)[0] for _ in range(n)]
    got = [generate_maze_map(1234, 7 + i, i % 5, g, 10, 12) for i in range(n)]
    ref_free = np.array([(c != 1).sum() for c in ref]); got_free = np.array([(c != 1).sum() for c, _ in got])
    assert abs(ref_free.mean() - got_free.mean()) < 4.0 and abs(ref_free.std() - got_free.std()) < 3.0
    marg_ref = np.mean([(c != 1) for c in ref], axis=0); marg_got = np.mean([(c != 1) for c, _ in got], axis=0)
    assert np.abs(marg_ref - marg_got).max() < 0.15
    thirsty = np.mean([(c == 3).sum() for c, _ in got]) / 10
    assert abs(thirsty - 0.7) < 0.05
    for c, (rx, ry) in got:
        assert c[rx, ry] == 0 and int(((c == 2) | (c == 3)).sum()) == 10
        free = c != 1
        seen = np.zeros_like(free); seen[rx, ry] = True
        queue = deque([(rx, ry)])
        while queue:
            x, y = queue.popleft()
            for dx, dy in ((-1, 0), (0, 1), (1, 0), (0, -1)):
                nx, ny = x + dx, y + dy
                if 0 <= nx < g and 0 <= ny < g and free[nx, ny] and not seen[nx, ny]:
                    seen[nx, ny] = True; queue.append((nx, ny))
        assert seen.sum() == free.sum()                                    # one connected system of rooms
    # a grid too small for a single room falls back to the cluster generator, like the fork (:463-467)
    small, _ = generate_maze_map(5, 0, 0, 6, 2, 3)
    assert (small != 1).sum() > 20


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line with the contract's
    keys, the reference arm's `impl` / `cpu_baseline` / zero-copy `e2e`, the same metric and workload naming
    as our arm.  (Timed sample shortened through PLANTOS_BENCH_REF_SECONDS.)"""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PLANTOS_BENCH_REF_SECONDS="1.5")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "1",
                          "--steps", "3", "--warmup", "3"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["steps"] == 3 and d["warmup"] == 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert "training preset" in d["config"]["workload"] and d["config"]["sample_envs"] == 64 * d["cpu_baseline"]["cores"]
