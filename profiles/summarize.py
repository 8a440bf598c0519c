"""Turn the ncu artefacts of profiles/prof_final.sh (read from gpurun_out/) into the tracked
summary files: profiles/<tag>_launches.csv, profiles/<tag>_ncu_summary.md, profiles/ncu_traffic.json."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
commit = sys.argv[2] if len(sys.argv) > 2 else None

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
]


def raw(rep):
    pre = rep[: -len(".ncu-rep")] + ".raw.csv"  # extracted on the GPU box (profiles/prof_r2b.sh); else read the report here
    if os.path.exists(pre):
        txt = open(pre).read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (vals[i], units[i]) for i, h in enumerate(hdr)}


def num(x):
    return float(x.replace(",", ""))


def to_bytes(val, unit):
    return num(val) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


script = sys.argv[3] if len(sys.argv) > 3 else "profiles/prof_r2.sh"
lines = [f"# ncu summary ({tag})", "", f"Commands: `{script}` (the default `python bench.py` command, `--clock-control none`; `predict_small_kernel`: "
         "the C5 forest on 1,048,576 rows of the shard).", ""]
src = os.path.join(OUT, "launchesF.csv")
if os.path.exists(src):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            agg[r[ki][:70]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w") as f:
        f.write(open(src).read())
    lines += ["## Launch list (gpu__time_duration.sum; cold-cache, serialised: shares only)", "", "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"| `{k}` | {len(v)} | {sum(v)/1e6:.3f} | {sum(v)/tot*100:.1f}% |")
    lines.append("")
traffic = {}
for rep, name in (("prof_growF", "grow_kernel"), ("prof_gatherF", "slot_gather_kernel"), ("prof_predictF", "predict_kernel"),
                  ("prof_smallF", "predict_small_kernel")):
    path = os.path.join(OUT, rep + ".ncu-rep")
    if not os.path.exists(path) and not os.path.exists(path[: -len(".ncu-rep")] + ".raw.csv"):
        continue
    d = raw(path)
    lines += [f"## `rfmc::{name}` (one launch, `--set full`)", "", "| metric | value |", "|---|---|"]
    for k in KEYS:
        if k in d:
            lines.append(f"| {k} | {d[k][0]} {d[k][1]} |")
    lines.append("")
    traffic[name] = {
        "capture": f"profiles/{tag}_ncu_summary.md (ncu --set full --clock-control none, one launch)", "commit": commit,
        "dram_bytes_per_launch": to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"]),
        # one shared-memory wavefront moves up to 128 bytes (the denominator of SURVEY 8d's secondary bound)
        "shared_wavefront_bytes_per_launch": num(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"][0]) * 128.0,
        "issue_slot_utilisation": num(d["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]) / 100.0,
        "ms_per_launch_under_ncu": num(d["gpu__time_duration.sum"][0]) * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(d["gpu__time_duration.sum"][1], 1.0),
    }
with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md"), "w") as f:
    f.write("\n".join(lines) + "\n")
workload = None
plain = os.path.join(OUT, "plainF.json")
if os.path.exists(plain):
    workload = json.load(open(plain))["config"]["workload"]
tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
old_doc = json.load(open(tj)) if os.path.exists(tj) else {}
kernels = {}
for name, k in old_doc.get("kernels", {}).items():  # kernels not captured this time keep their own capture label
    kernels[name] = dict(k)
    kernels[name].setdefault("capture", old_doc.get("capture"))
    kernels[name].setdefault("commit", old_doc.get("commit"))
kernels.update(traffic)
with open(tj, "w") as f:
    json.dump({"workload": workload or old_doc.get("workload"), "kernels": kernels}, f, indent=1)
print("\n".join(lines))
