#!/bin/bash
python __graft_entry__.py smoke 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -x -q -k "maze or set_state or push_maps or abi" 2>&1 | tail -3
python - <<'PY'
import sys, torch
sys.argv=['x','96']
exec(open('tools/bench_presets.py').read().split("steps = int")[0])
T = PRESETS["training"]
run("training preset, map_source='maze' (device maze generator), 131072 envs", 131072, 160, map_source="maze", **T)
PY
