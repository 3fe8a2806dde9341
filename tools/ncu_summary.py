"""Summarise an .ncu-rep (ncu --set full) into the text format kept under profiles/: one line per metric, one column per
launch.  Usage: python tools/ncu_summary.py report.ncu-rep "header text" [first_launch [n_launches]] > profiles/xyz.txt"""
import csv, io, subprocess, sys

METRICS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size",
           "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active"]
rep, header = sys.argv[1], sys.argv[2]
first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
count = int(sys.argv[4]) if len(sys.argv) > 4 else 3
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:][first:first + count]
print(header)
for m in METRICS:
    if m in hdr:
        i = hdr.index(m)
        print(f"{m} [{units[i]}]: {[r[i][:60] for r in data]}")
