"""Driver for the `ncu --set full` capture of the kernels in bench.py's per-kernel roofline table: launches each of
them at its benchmark shape (B = 256) a few times.  Run under ncu (see tools/gpu_ncu_top.sh); the summary and the
per-launch DRAM traffic go to profiles/r02_ncu_top_kernels.txt and profiles/r02_ncu_traffic.json (tools/ncu_traffic.py)."""
import sys
sys.path.insert(0, ".")
import torch
import bench
import m3l_b200  # noqa: F401

peaks = bench.load_peaks()
# time_kernel_cold launches each kernel 2 + 8 times; under ncu only the launches selected with --launch-skip/-c count
bench.time_kernel_cold.__defaults__ = (1,)
rows = bench.per_kernel_roofline(peaks, 2.4)
torch.cuda.synchronize()
for r in rows:
    print(r["kernel"], r["us_per_launch"])
