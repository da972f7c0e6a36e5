"""profiles/r02_ncu_top_kernels.txt + profiles/r02_ncu_traffic.json from an .ncu-rep captured with tools/ncu_top_kernels.py:
one line per profiled launch (time, DRAM read / write bytes, L2 bytes, tensor / issue utilisation) and, per roofline-table
key, dram__bytes_read.sum + dram__bytes_write.sum of the LAST launch of that kernel (warm-up launches come first).
usage: python tools/ncu_traffic.py report.ncu-rep"""
import csv, json, re, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {n: hdr.index(n) for n in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                  "lts__t_bytes.sum", "launch__grid_size") if n in hdr}
opt = {"tensor%": "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
       "issue%": "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "dram%": "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts%": "lts__throughput.avg.pct_of_peak_sustained_elapsed"}
opt = {k: hdr.index(v) for k, v in opt.items() if v in hdr}


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


lines, seq = [], []
for r in rows[2:]:
    nm = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("m3l::<unnamed>::", "")
    rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    t = r[col["gpu__time_duration.sum"]] + units[col["gpu__time_duration.sum"]]
    extra = " ".join(f"{k}={r[i]}" for k, i in opt.items())
    lines.append(f"{nm[:60]:60s} time={t:>12s} dram_rd={rd / 1e6:9.2f}MB dram_wr={wr / 1e6:9.2f}MB grid={r[col['launch__grid_size']]} {extra}")
    seq.append((nm, rd + wr))
open("profiles/r02_ncu_top_kernels.txt", "w").write("\n".join(lines) + "\n")
# the driver (tools/ncu_top_kernels.py) launches the table's kernels in this order, 2 warm-up launches + 1 each: the
# DRAM bytes of the LAST launch of every group of three is what bench.py reports as `traffic`
order = ["wgrad_ff", "wgrad_ff2", "wgrad_qkv", "ln_mlp_fwd_save", "attn_bwd", "attn_fwd", "dgrad_ff2", "dgrad_ff1_ln", "qkv_fwd",
         "out_proj", "ln_bwd", "ln_fwd", "dgrad_qkv_ln", "out_proj_dgrad"]
if len(seq) == 3 * len(order) + 1 and "attn_fwd" in seq[0][0]:      # the set-up launch that produces O / LSE for the backward
    seq = seq[1:]
assert len(seq) == 3 * len(order), (len(seq), len(order))
traffic = {k: seq[3 * i + 2][1] for i, k in enumerate(order)}
json.dump(traffic, open("profiles/r02_ncu_traffic.json", "w"), indent=1)
json.dump({"_launch_sequence": [[n, b] for n, b in seq], "_order": order}, open("profiles/r02_ncu_traffic_raw.json", "w"), indent=1)
print("\n".join(lines))
