"""Per-kernel table from an `ncu --page raw --csv` export (the format of profiles/*_all_kernels_ncu_summary.txt).

    ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats \\
        --section SchedulerStats --section ComputeWorkloadAnalysis --clock-control none -o /tmp/all \\
        python scripts/profile_all_kernels.py
    ncu -i /tmp/all.ncu-rep --page raw --csv > gpurun_out/all_kernels_raw.csv
    python scripts/summarise_ncu_raw.py gpurun_out/all_kernels_raw.csv [hbm_peak_gbs] > profiles/rNN_all_kernels_ncu_summary.txt
"""
import csv
import json
import os
import sys

path = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
if peak is None:
    try:
        peak = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6571.9
rows = list(csv.reader(open(path)))
hdr, units, body = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def num(r, name, scale_unit=None):
    v = r[col[name]].replace(",", "")
    try:
        x = float(v)
    except ValueError:
        return float("nan")
    u = units[col[name]]
    if scale_unit == "us":
        x *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    if scale_unit == "GB/s":
        x *= {"byte/s": 1e-9, "Kbyte/s": 1e-6, "Mbyte/s": 1e-3, "Gbyte/s": 1.0, "Tbyte/s": 1e3}.get(u, 1.0)
    return x


last = {}
for r in body:
    key = (r[col["Kernel Name"]], r[col["Grid Size"]], r[col["Block Size"]])
    last[key] = r  # last launch of each (kernel, grid, block)
print(f"# ncu per-kernel summary: sections SpeedOfLight / MemoryWorkloadAnalysis / Occupancy / LaunchStats / SchedulerStats / "
      f"ComputeWorkloadAnalysis, --clock-control none")
print(f"# produced by scripts/profile_all_kernels.py under ncu, summarised by scripts/summarise_ncu_raw.py; last launch of each "
      f"(kernel, grid, block).  dram_GBps = dram__bytes.sum.per_second; HBM copy peak (MEASURED_PEAKS.json) = {peak} GB/s")
print("kernel | grid | block | us | dram_GBps | frac_of_copy_peak | dram_active% | sm_throughput% | regs | achieved_occ% | fp64_pipe% | issue_active%")
for (name, grid, block), r in last.items():
    gbs = num(r, "dram__bytes.sum.per_second", "GB/s")
    print(" | ".join([name[:64], grid, block, f"{num(r, 'gpu__time_duration.sum', 'us'):.1f}", f"{gbs:.0f}", f"{gbs / peak:.3f}",
                      f"{num(r, 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f}",
                      f"{num(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f}",
                      f"{num(r, 'launch__registers_per_thread'):.0f}",
                      f"{num(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f}",
                      f"{num(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):.1f}",
                      f"{num(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f}"]))
