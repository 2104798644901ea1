"""Summarise an .ncu-rep (raw page) into the handful of metrics quoted in profiles/."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[0]; units = rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("----")
    for k in want:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k} = {r[i][:100]} {units[i]}")
