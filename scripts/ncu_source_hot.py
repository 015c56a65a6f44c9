"""ncu_source_hot.py REPORT KERNEL_REGEX [TOP]: per-source-line hot spots of one kernel from an .ncu-rep captured with
--import-source on (-lineinfo build): stall samples, warp instructions, thread efficiency, L1/L2 sectors."""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, agg = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] in ("Function Name",) or not r[0].isdigit() or r[2] != "-":
        continue
    d = dict(zip(hdr, r))
    key = (cur_file, int(r[0]))
    samples = int(d.get("# Samples") or 0)
    inst = int(d.get("Instructions Executed") or 0)
    tinst = int(d.get("Thread Instructions Executed") or 0)
    a = agg.setdefault(key, [0, 0, 0, r[1]])
    a[0] += samples
    a[1] += inst
    a[2] += tinst
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
print("%-18s %7s %7s %6s  %s" % ("file:line", "stall%", "inst%", "thr/w", "source"))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-18s %6.2f%% %6.2f%% %6.1f  %s" % ("%s:%d" % key, 100.0 * a[0] / tot_s, 100.0 * a[1] / tot_i, a[2] / max(a[1], 1), a[3].strip()[:110]))
