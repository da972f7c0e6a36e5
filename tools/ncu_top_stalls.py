"""Prints the top stalled SASS instructions per kernel from `ncu --page source --csv` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for sec in sections:
    hdr = sec["hdr"]; idx = {h: i for i, h in enumerate(hdr)}
    data = sec["data"]
    tot = sum(int(r[idx["# Samples"]]) for r in data) or 1
    print("==", sec["name"][:100], "samples", tot)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(int(r[idx[h]]) for r in data) for h in stall_cols}
    print("   overall:", sorted(((k, round(100 * v / tot, 1)) for k, v in agg.items() if v), key=lambda kv: -kv[1])[:6])
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:topn]:
        s = int(r[idx["# Samples"]])
        st = sorted(((h, int(r[idx[h]])) for h in stall_cols if int(r[idx[h]]) > 0), key=lambda kv: -kv[1])[:3]
        print(f"{100 * s / tot:5.1f}% {r[idx['Source']].strip()[:66]:66s} {st}")
