"""Turns the ncu artefacts a gpurun call left in gpurun_out/ into the small text summaries kept under profiles/.
   python tools/summarize_profiles.py <round tag> <launch list csv> [<name>=<ncu-rep> ...]"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches = sys.argv[1], sys.argv[2]
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

# ---- launch list: per-kernel totals and the kernels' share of the step ----
with open(launches) as f:
    lines = [l for l in f if l.startswith('"')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0].replace("void ", "")
    t = float(row["Metric Value"].replace(",", "")) / 1000.0
    if t < 4 and "auction" in name and "init" not in name and "finalize" not in name:
        name += " [returns at once: not its turn]"
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(v[1] for v in agg.values())
with open(os.path.join(out, f"{tag}_bench_launches_summary.txt"), "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, {os.path.basename(launches)}\n")
    f.write("# per-launch times are cold-cache and serialised: read the SHARES, not the absolutes\n")
    f.write(f"{'kernel':78s} {'launches':>8s} {'total us':>10s} {'avg us':>8s} {'share':>6s}\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k[:78]:78s} {v[0]:8d} {v[1]:10.1f} {v[1] / v[0]:8.1f} {v[1] / tot:6.3f}\n")
    f.write(f"{'total':78s} {sum(v[0] for v in agg.values()):8d} {tot:10.1f}\n")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
traffic = {}
tpath = os.path.join(out, "traffic.json")
if os.path.exists(tpath):
    traffic = json.load(open(tpath))
for spec in sys.argv[3:]:
    name, rep = spec.split("=")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, val = rows[0], rows[1], rows[2]
    m = {h: (val[i], units[i]) for i, h in enumerate(hdr)}
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    srows = [r for r in csv.reader(src.splitlines()) if len(r) > 8 and r[2] == "-" and r[4].isdigit()]
    ti, ts = sum(int(r[7]) for r in srows), sum(int(r[4]) for r in srows)
    with open(os.path.join(out, f"{tag}_{name}_ncu.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on, one launch of {m.get('Kernel Name', ('?',))[0]}\n")
        f.write(f"# report: {os.path.basename(rep)} (not committed: binary); workload tools/profile_iter.py (1 M x 512, K=128)\n")
        for w in WANT:
            if w in m:
                f.write(f"{w:75s} {m[w][0]:>16s} {m[w][1]}\n")
        for h in hdr:
            if h.startswith("smsp__warp_issue_stalled") and h.endswith("per_warp_active.pct"):
                try:
                    if float(m[h][0]) >= 2.0:
                        f.write(f"{h:75s} {m[h][0]:>16s} %\n")
                except ValueError:
                    pass
        f.write(f"\n# source lines by warp-stall samples (total instructions {ti}, samples {ts})\n")
        for r in sorted(srows, key=lambda r: -int(r[4]))[:24]:
            f.write(f"L{r[0]:>5s} inst {int(r[7]) / max(ti, 1):6.1%} samples {int(r[4]) / max(ts, 1):6.1%} | {r[1].strip()[:110]}\n")
    try:
        def to_bytes(x):
            v, u = float(x[0].replace(",", "")), x[1].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        kname = m["Kernel Name"][0].split("(")[0].replace("rqk::", "").replace("void ", "").strip()
        traffic[kname] = to_bytes(m["dram__bytes_read.sum"]) + to_bytes(m["dram__bytes_write.sum"])
    except Exception as e:
        print("traffic:", e)
json.dump(traffic, open(tpath, "w"), indent=1)
print(open(os.path.join(out, f"{tag}_bench_launches_summary.txt")).read())
print(traffic)
