"""Print the hottest SASS lines (warp-stall samples) of one kernel in an .ncu-rep.
usage: python tools/ncu_hot.py report.ncu-rep kernel_regex [top]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = 0
while start < len(rows):
    if rows[start] and rows[start][0] == "Kernel Name":
        name = rows[start][1]
        hdr = rows[start + 1]
        end = start + 2
        while end < len(rows) and not (rows[end] and rows[end][0] == "Kernel Name"):
            end += 1
        data = rows[start + 2:end]
        si, src = hdr.index("# Samples"), hdr.index("Source")
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[si]) for r in data)
        print("==", name[:100], "samples", tot, "sass lines", len(data))
        agg = {}
        for r in data:
            for c in stall_cols:
                agg[hdr[c]] = agg.get(hdr[c], 0) + int(r[c])
        print("  stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
        top = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:top_n]
        for i in sorted(top):
            r = data[i]
            st = sorted([(int(r[c]), hdr[c][6:]) for c in stall_cols], reverse=True)[:2]
            print(f"  {i:5d} {int(r[si]):6d} {100*int(r[si])/max(tot,1):5.1f}%  {r[src].strip()[:80]:80s} {st}")
        break   # first matching launch only
    start += 1
