"""Per-launch summary of an .ncu-rep: time, DRAM bytes, L2 traffic, tensor/issue utilisation.
usage: python tools/ncu_summary.py report.ncu-rep [kernel_regex]"""
import csv, subprocess, sys, re
rep = sys.argv[1]
cmd = ["ncu", "-i", rep, "--page", "raw", "--csv"]
if len(sys.argv) > 2:
    cmd += ["--kernel-name", f"regex:{sys.argv[2]}"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = [("gpu__time_duration.sum", "time"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("lts__t_bytes.sum", "l2_bytes"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"),
        ("smsp__inst_executed.sum", "inst")]
idx = {}
for name, short in want:
    if name in hdr:
        idx[short] = hdr.index(name)
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    nm = re.sub(r"\(.*", "", r[ki]).replace("m3l::<unnamed>::", "")[:48]
    parts = [f"{nm:48s}"]
    for short, i in idx.items():
        parts.append(f"{short}={r[i]}{units[i] if short in ('time','dram_rd','dram_wr','l2_bytes') else ''}")
    print(" ".join(parts))
