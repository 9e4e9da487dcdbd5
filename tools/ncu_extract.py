"""Summarise an .ncu-rep (ncu -i <rep> --page raw --csv): per kernel duration, DRAM bytes, tensor-pipe / DRAM utilisation."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
idx = {h: i for i, h in enumerate(hdr)}
units = rows[1]
for r in rows[2:]:
    print("==", r[idx["Kernel Name"]][:90])
    for k in want[1:]:
        if k in idx:
            print(f"   {k:80s} {r[idx[k]]:>16s} {units[idx[k]]}")
