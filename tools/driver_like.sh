#!/bin/bash
set -x
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 | cut -c1-400
python bench.py --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('metric','value','n_gpus','steps','warmup','ms_per_step','gpu_launches')}); print(d['roofline']); print(d['e2e']['value'], d['cpu_baseline']); print(d['clocks']); print(d['config']['launch'])"
