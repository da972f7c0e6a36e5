"""Aggregate ncu warp-stall samples of one kernel per CUDA source line.
usage: python tools/ncu_lines.py report.ncu-rep object.o kernel_regex mangled_substring [top]
The SASS -> line map comes from `nvdisasm -g` on the cubin inside object.o (same build!)."""
import csv, re, subprocess, sys, tempfile, os, glob
rep, obj, kern, mangled = sys.argv[1:5]
top_n = int(sys.argv[5]) if len(sys.argv) > 5 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = glob.glob(tmp + "/*.cubin")[0]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# locate function
lines = []  # per instruction: (file:line chain)
infn = False
cur = "?"
for l in sass:
    if l.startswith(".text."):
        infn = mangled in l
        cur = "?"
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = f"{os.path.basename(m.group(1))}:{m.group(2)}"
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        if inl:
            cur += " <- " + " <- ".join(f"{os.path.basename(a)}:{b}" for a, b in inl)
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name")
hdr = rows[start + 1]
end = start + 2
while end < len(rows) and not (rows[end] and rows[end][0] == "Kernel Name"):
    end += 1
data = rows[start + 2:end]
si = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
print("kernel:", rows[start][1][:110], "| sass", len(data), "| mapped", len(lines))
agg = {}
for i, r in enumerate(data):
    key = lines[i] if i < len(lines) else "?"
    # outermost (non-inlined) location = last element of chain
    outer = key.split(" <- ")[-1]
    a = agg.setdefault(outer, [0, {}, {}])
    a[0] += int(r[si])
    for c in stall_cols:
        v = int(r[c])
        if v:
            a[1][hdr[c][6:]] = a[1].get(hdr[c][6:], 0) + v
    inner = key.split(" <- ")[0]
    a[2][inner] = a[2].get(inner, 0) + int(r[si])
tot = sum(a[0] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    st = sorted(a[1].items(), key=lambda kv: -kv[1])[:3]
    inner = sorted(a[2].items(), key=lambda kv: -kv[1])[:2]
    print(f"{a[0]:6d} {100*a[0]/tot:5.1f}%  {k:24s} {st}  inner={inner}")
