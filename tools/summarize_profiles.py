#!/usr/bin/env python3
"""Turn the raw evidence a GPU run left in gpurun_out/ev/ into the committed summaries under profiles/.
usage: tools/summarize_profiles.py <round-tag, e.g. r1> [ev-dir]
Inputs (see profiles/<tag>_summary.md for the commands that produce them):
  launches.csv        ncu --metrics gpu__time_duration.sum --clock-control none (launch list)
  prof_fast_*.ncu-rep ncu --set full of one k_step_fast launch
  steady_dram.csv     ncu --replay-mode application --cache-control none, dram bytes of 4 launches
  bench_*.json, pytest_gpu.log"""
import collections, csv, glob, io, json, os, shutil, subprocess, sys
tag = sys.argv[1]
ev = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/ev"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")

def csv_rows(path):
    lines = [l for l in open(path) if l.startswith('"')]
    return list(csv.DictReader(io.StringIO("".join(lines))))

# launch list
rows = csv_rows(os.path.join(ev, "launches.csv"))
shutil.copy(os.path.join(ev, "launches.csv"), os.path.join(out, f"{tag}_launches.csv"))
agg = collections.OrderedDict()
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum": continue
    ns = float(r["Metric Value"]) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r["Metric Unit"], 1)
    a = agg.setdefault(r["Kernel Name"][:72], [0, 0.0]); a[0] += 1; a[1] += ns
tot = sum(a[1] for a in agg.values())
summ = [{"kernel": k, "launches": a[0], "mean_us": round(a[1] / a[0] / 1e3, 2), "total_us": round(a[1] / 1e3, 1),
         "share": round(a[1] / tot, 4)} for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]
json.dump(summ, open(os.path.join(out, f"{tag}_launches_summary.json"), "w"), indent=1)

# full-set metrics of the hot kernel
rep = sorted(glob.glob(os.path.join(ev, "prof_fast_*.ncu-rep")))[-1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
names, units, vals = rr[0], rr[1], rr[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
sel = {}
with open(os.path.join(out, f"{tag}_fast_kernel_metrics.csv"), "w") as f:
    f.write("metric,unit,value\n")
    for i, n in enumerate(names):
        if n in want or (n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio")):
            f.write(f"{n},{units[i]},{vals[i]}\n"); sel[n] = (units[i], float(vals[i].replace(",", "")))
def to_bytes(n):
    u, v = sel[n]; return int(v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u])

# steady-state DRAM
st = csv_rows(os.path.join(ev, "steady_dram.csv"))
shutil.copy(os.path.join(ev, "steady_dram.csv"), os.path.join(out, f"{tag}_steady_state_dram.csv"))
rd = [float(r["Metric Value"]) for r in st if r["Metric Name"] == "dram__bytes_read.sum"]
wr = [float(r["Metric Value"]) for r in st if r["Metric Name"] == "dram__bytes_write.sum"]
cold_r, cold_w = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
traffic = {
    "kernel": rows and [r["Kernel Name"] for r in rows if "k_step_fast" in r["Kernel Name"]][0][:60],
    "workload": "131072 envs, training preset (D=107)",
    "source": f"ncu --set full --clock-control none, {os.path.basename(rep)} (caches flushed between replay passes)",
    "dram_bytes_read": cold_r, "dram_bytes_write": cold_w, "dram_bytes_per_launch": cold_r + cold_w,
    "note": "cold-cache per-launch figure: ncu flushes L2 before each pass, so state reads come from DRAM while most of "
            "the bytes written are still dirty in L2 when the kernel ends. steady_state_* = mean of 4 consecutive launches "
            "inside the bench loop (ncu --replay-mode application --cache-control none, profiles/%s_steady_state_dram.csv)." % tag,
    "steady_state_dram_bytes_read": int(sum(rd) / len(rd)), "steady_state_dram_bytes_write": int(sum(wr) / len(wr)),
    "steady_state_dram_bytes_per_launch": int(sum(rd) / len(rd) + sum(wr) / len(wr)),
    "algorithmic_bytes_per_launch": 131072 * 441,
}
json.dump(traffic, open(os.path.join(out, f"{tag}_traffic.json"), "w"), indent=1)
for src, dst in [("bench_1gpu.json", f"{tag}_bench_1gpu.json"), ("bench_2gpu.json", f"{tag}_bench_2gpu.json"),
                 ("bench_8gpu.json", f"{tag}_bench_8gpu.json"), ("pytest_gpu.log", f"{tag}_pytest_gpu.log"),
                 ("sizes.log", f"{tag}_sizes.log")]:
    if os.path.exists(os.path.join(ev, src)): shutil.copy(os.path.join(ev, src), os.path.join(out, dst))
print(json.dumps(summ[:4], indent=1)); print(json.dumps(traffic, indent=1))
