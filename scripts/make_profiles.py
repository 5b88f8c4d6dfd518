#!/usr/bin/env python
"""Condenses what a gpurun profiling call left in gpurun_out/ into the tracked profiles/ directory:
launch list (ncu --metrics gpu__time_duration.sum), the headline metrics of the `ncu --set full`
captures (read with `ncu -i ... --page raw --csv`), and the bench lines of the same code.
usage: scripts/make_profiles.py <tag> [round prefix] (files gpurun_out/*<tag>*), e.g. r2f r2"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1]
RND = sys.argv[2] if len(sys.argv) > 2 else "r2"

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def launches():
    lines = open(os.path.join(OUT, f"launches{tag}.csv")).read().split("\n")
    i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(lines[i:]))
    seq, agg = [], collections.OrderedDict()
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        us = float(r["Metric Value"].replace(",", "")) / 1000.0
        seq.append(f"{int(r['ID']):4d} {name:44s} {us:9.2f} us  grid {r['Grid Size']} block {r['Block Size']}")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    with open(os.path.join(PROF, f"{RND}_launch_sequence.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400; "
                "bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-zipf --query-records 20000000\n"
                "# (cold-cache, serialised launches: compare shares, not absolutes)\n")
        f.write("\n".join(seq) + "\n")
    with open(os.path.join(PROF, f"{RND}_launches_summary.csv"), "w") as f:
        f.write("kernel,launches,total_us,mean_us,share\n")
        for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{name}\",{n},{us:.1f},{us / n:.2f},{us / total:.3f}\n")


def capture(name):
    rep = os.path.join(OUT, f"prof{tag}_{name}.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.split("\n")))
    h, units, v = r[0], r[1], r[2]
    out = {"kernel": v[h.index("Kernel Name")]}
    with open(os.path.join(PROF, f"{RND}_ncu_{name}.csv"), "w") as f:
        f.write("metric,unit,value\n")
        f.write(f"kernel,,\"{out['kernel']}\"\n")
        for k in KEEP:
            if k in h:
                j = h.index(k)
                f.write(f"{k},{units[j]},{v[j]}\n")
                out[k] = (v[j], units[j])
    return out


def main():
    launches()
    traffic = {}
    for name, kern in (("merge", "k_merge_stage"), ("scan", "k_index_scan"), ("build", "k_index_build"), ("hot", "k_merge_hot")):
        if not os.path.exists(os.path.join(OUT, f"prof{tag}_{name}.ncu-rep")):
            continue
        m = capture(name)
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        rd = float(m["dram__bytes_read.sum"][0]) * scale[m["dram__bytes_read.sum"][1]]
        wr = float(m["dram__bytes_write.sum"][0]) * scale[m["dram__bytes_write.sum"][1]]
        traffic[kern] = {"bytes_per_launch": rd + wr, "read": rd, "write": wr,
                         "us_under_ncu": float(m["gpu__time_duration.sum"][0]),
                         "source": f"profiles/{RND}_ncu_{name}.csv (ncu --set full --clock-control none, one launch)"}
    with open(os.path.join(PROF, "traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    for src, dst in ((f"bench{tag}.json", f"{RND}_bench_n1.json"), (f"bench{tag}_ref.json", f"{RND}_bench_n1_reference.json"),
                     (f"bench{tag}_mesh.json", f"{RND}_bench_mesh_n1.json")):
        if os.path.exists(os.path.join(OUT, src)):
            shutil.copy(os.path.join(OUT, src), os.path.join(PROF, dst))
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
