"""Turns the ncu artefacts of a profiling run (gpurun_out/) into the committed text summaries under profiles/.
usage: python scripts/ncu_summaries.py <launches.csv> <report.ncu-rep> <tag>"""
import collections
import csv
import json
import subprocess
import sys

launches, rep, tag = sys.argv[1:4]
rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
cols = rows[0]
ki, vi = cols.index("Kernel Name"), cols.index("Metric Value")
data = rows[1:]
names = [r[ki] for r in data]
vals = [float(r[vi].replace(",", "")) for r in data]
last = max(i for i, n in enumerate(names) if "k_extent_partial" in n)
agg = collections.OrderedDict()
for n, v in zip(names[last:], vals[last:]):
    key = n.split("(")[0].replace("void ", "")
    if "cub" in key:
        key = "cub::" + (key.split("::")[2][:40] if key.count("::") >= 2 else key[:50])
    a = agg.setdefault(key, [0, 0])
    a[0] += v
    a[1] += 1
tot = sum(v[0] for v in agg.values())
out = ["# ncu launch list, last (timed) step of: python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline",
       "# gpu__time_duration.sum, --clock-control none; per-launch times are cold-cache and serialised: compare SHARES",
       "# launches in the step: %d, total %.3f ms" % (len(names) - last, tot / 1e6),
       "%-50s %5s %12s %7s" % ("kernel", "count", "ns", "share")]
for k, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    out.append("%-50s %5d %12.0f %6.1f%%" % (k[:50], c, v, 100 * v / tot))
open("profiles/%s_launches_summary.txt" % tag, "w").write("\n".join(out) + "\n")
print("\n".join(out[:12]))

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum",
        "sm__cycles_elapsed.max"]
idx = {h: i for i, h in enumerate(hdr)}
txt = ["# ncu --set full --clock-control none --import-source on, one warm step of the 1.6M-gas disc (bench.py --steps 1 --warmup 3)"]
traffic = {}
for r in rows[2:]:
    kn = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
    txt.append("---- " + kn)
    for w in want:
        if w in idx:
            txt.append("  %-72s %s %s" % (w, r[idx[w]], units[idx[w]]))
    def tobytes(v, u):
        f = float(v.replace(",", ""))
        return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    rd = tobytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
    wr = tobytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
    traffic[kn.split("<")[0]] = rd + wr
open("profiles/%s_ncu_full_summary.txt" % tag, "w").write("\n".join(txt) + "\n")
json.dump(dict(source="ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, 1.6M-gas disc, one warm step",
               bytes_per_launch=traffic), open("profiles/%s_traffic.json" % tag, "w"), indent=1)
print("\n".join(txt))
