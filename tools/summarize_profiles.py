#!/usr/bin/env python3
"""Turn the raw evidence of tools/gpu_evidence.sh (gpurun_out/ev2/) into the committed summaries under profiles/.
usage: tools/summarize_profiles.py <round-tag, e.g. r2> [ev-dir]"""
import collections, csv, glob, io, json, os, shutil, subprocess, sys
tag = sys.argv[1]
ev = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/ev2"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")


def csv_rows(path):
    lines = [l for l in open(path) if l.startswith('"')]
    return list(csv.DictReader(io.StringIO("".join(lines))))


def last_json(path):
    try:
        return json.loads(open(path).read().strip().splitlines()[-1])
    except Exception:
        return None


# ---- bench lines and logs
def last_json(path):      # (multi-rank runs: other text may share the file)
    try:
        return json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])
    except Exception:
        return None


for name in ("bench_1gpu", "bench_1gpu_20steps", "bench_reference_cpu", "bench_nostagger", "bench_preset_default", "bench_preset_xl",
             "bench_4096", "bench_2gpu", "bench_4gpu", "bench_8gpu", "bench_2gpu_640steps", "bench_4gpu_640steps", "bench_8gpu_640steps"):
    d = last_json(os.path.join(ev, name + ".json"))
    if d is not None:
        json.dump(d, open(os.path.join(out, f"{tag}_{name}.json"), "w"))
scale = {}
for n in (1, 2, 4, 8):
    for suffix, key in (("", "steps20"), ("_640steps", "steps640")):
        d = last_json(os.path.join(ev, ("bench_1gpu_20steps" if suffix == "" else "bench_1gpu") + ".json")) if n == 1 else \
            last_json(os.path.join(ev, f"bench_{n}gpu{suffix}.json"))
        if d is not None:
            scale.setdefault(str(n), {})[key if n > 1 or suffix == "" else "steps3000"] = {
                "steps": d["steps"], "us_per_step": round(d["ms_per_step"] * 1e3, 2), "value": d["value"],
                "frac": round(d["roofline"]["frac"], 3), "e2e": d["e2e"]["value"],
                "e2e_ms_per_step_by_rank": d["e2e"].get("ms_per_step_by_rank"),
                "stats_allreduces_in_timed_window": d["config"].get("stats_allreduces_in_timed_window")}
json.dump(scale, open(os.path.join(out, f"{tag}_scaling.json"), "w"), indent=1)
for n in (2, 4, 8):
    if os.path.exists(os.path.join(ev, f"topo_{n}gpu.txt")):
        shutil.copy(os.path.join(ev, f"topo_{n}gpu.txt"), os.path.join(out, f"{tag}_topo_{n}gpu.txt"))
    err = os.path.join(ev, f"bench_{n}gpu.err")
    if os.path.exists(err):     # the NCCL lines the driver's rank check looks for
        keep = [l for l in open(err, errors="replace") if "NCCL INFO" in l and ("nranks" in l or "Init COMPLETE" in l or "NVLS" in l)]
        open(os.path.join(out, f"{tag}_nccl_{n}gpu.log"), "w").write("".join(keep[:80]))
loops = {}
for name in ("bench_loop_graph", "bench_loop_graph_plain", "bench_loop_eager", "bench_loop_eager_plain", "bench_loop_graph_nostagger"):
    d = last_json(os.path.join(ev, name + ".json"))
    if d is not None:
        loops[name] = {"us_per_step": round(d["ms_per_step"] * 1e3, 2), "frac": round(d["roofline"]["frac"], 3),
                       "launch": d["config"]["launch"], "episodes": d["episode_stats"]["episodes"]}
json.dump(loops, open(os.path.join(out, f"{tag}_bench_loops.json"), "w"), indent=1)
for name in ("presets.log", "pytest_gpu.log", "gpu.txt"):
    if os.path.exists(os.path.join(ev, name)):
        text = open(os.path.join(ev, name)).read()
        if name == "pytest_gpu.log":
            text = "\n".join(text.strip().splitlines()[-3:]) + "\n"
        open(os.path.join(out, f"{tag}_{name}"), "w").write(text)

# ---- launch list (no cache flush between launches)
rows = csv_rows(os.path.join(ev, "launches.csv"))
shutil.copy(os.path.join(ev, "launches.csv"), os.path.join(out, f"{tag}_launches.csv"))
agg = collections.OrderedDict()
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    ns = float(r["Metric Value"]) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r["Metric Unit"], 1)
    a = agg.setdefault(r["Kernel Name"][:80], [0, 0.0]); a[0] += 1; a[1] += ns
tot = sum(a[1] for a in agg.values())
summ = [{"kernel": k, "launches": a[0], "mean_us": round(a[1] / a[0] / 1e3, 2), "total_us": round(a[1] / 1e3, 1),
         "share": round(a[1] / tot, 4)} for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]
json.dump(summ, open(os.path.join(out, f"{tag}_launches_summary.json"), "w"), indent=1)

# ---- full-set metrics of the two hot kernels
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "lts__t_sector_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
for which in ("rollout", "step"):
    rep = os.path.join(ev, f"prof_{which}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    names, units, vals = rr[0], rr[1], rr[2]
    with open(os.path.join(out, f"{tag}_{which}_kernel_metrics.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        w.writerow(["kernel", "", vals[names.index("Kernel Name")] if "Kernel Name" in names else ""])
        for i, n in enumerate(names):
            if n in want or n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio"):
                w.writerow([n, units[i], vals[i]])

# ---- the kernels off the default loop (tools/run_other_kernels.py): one column of metrics per captured launch
rep = os.path.join(ev, "prof_others.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    names, units = rr[0], rr[1]
    seen, cols = set(), []
    for vals in rr[2:]:
        k = vals[names.index("Kernel Name")]
        g = vals[names.index("launch__grid_size")] if "launch__grid_size" in names else ""
        if (k, g) in seen:
            continue
        seen.add((k, g)); cols.append(vals)
    with open(os.path.join(out, f"{tag}_other_kernels_metrics.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [c[names.index("Kernel Name")][:60] for c in cols])
        for i, n in enumerate(names):
            if n in want or n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio"):
                w.writerow([n, units[i]] + [c[i] for c in cols])

# ---- steady-state DRAM traffic
import hashlib
def kernel_source_sha():
    # written by tools/gpu_evidence.sh at CAPTURE time (both DRAM stages must agree); falls back to the current tree
    shas = set()
    for st in ("dram_rollout", "dram_step"):
        f = os.path.join(ev, f"kernel_source_sha_{st}.txt")
        if os.path.exists(f):
            shas.add(open(f).read().strip())
    if len(shas) == 1:
        return shas.pop()
    if len(shas) > 1:
        return "mixed"
    sys.path.insert(0, root)
    from rl_env_b200.build import kernel_source_hash
    return kernel_source_hash()
traffic = {"envs": 131072, "preset": "training", "kernel_source_sha": kernel_source_sha(), "how": "ncu --replay-mode application --cache-control none (no cache flush), bench.py timed loop", "kernels": {}}
for which, kname, steps in (("rollout", "k_rollout_tile", 16), ("step", "k_step_tile", 1)):
    path = os.path.join(ev, f"steady_dram_{which}.csv")
    if not os.path.exists(path):
        continue
    rows = csv_rows(path)
    shutil.copy(path, os.path.join(out, f"{tag}_steady_dram_{which}.csv"))
    rd = [float(r["Metric Value"]) for r in rows if r["Metric Name"] == "dram__bytes_read.sum"]
    wr = [float(r["Metric Value"]) for r in rows if r["Metric Name"] == "dram__bytes_write.sum"]
    if rd and wr:
        traffic["kernels"][kname] = {"dram_bytes_per_step": int((sum(rd) / len(rd) + sum(wr) / len(wr)) / steps),
                                     "read_per_step": int(sum(rd) / len(rd) / steps), "write_per_step": int(sum(wr) / len(wr) / steps),
                                     "launches_sampled": len(rd), "steps_per_launch": steps,
                                     "source": f"profiles/{tag}_steady_dram_{which}.csv"}
json.dump(traffic, open(os.path.join(out, f"{tag}_traffic.json"), "w"), indent=1)
print(json.dumps(traffic["kernels"], indent=1))
print(json.dumps(loops, indent=1))
print(json.dumps(summ[:6], indent=1))
